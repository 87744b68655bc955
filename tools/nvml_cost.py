import sys, time, statistics
sys.path.insert(0, '.')
import torch, pynvml
from pcq_import import pcq
S, B = pcq.synth, pcq.binding
ctx = pcq.Context(0)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
sp = S.uniform_spec(1 << 27, B.LAYOUT_LAS, 1)
buf = torch.empty(sp.n_points * sp.record_len + 256, dtype=torch.uint8, device="cuda:0")
mm, desc = S.device_points(ctx, sp, buf.data_ptr())
df = pcq.DeviceFile.wrap(ctx, desc, buf.data_ptr(), keepalive=buf)
s = pcq.BoundsSearcher((0, 0, 0), (10000.0, 10000.0, 5000.0))
c = pcq.CountCollector(ctx)
impl = pcq.SearchImplementation.Optimized
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
def launches(n):
    t0 = time.perf_counter()
    for _ in range(n): s.search_files([df], impl, [c])
    t1 = time.perf_counter(); ctx.synchronize(); return (t1 - t0) / n * 1e6
print("enqueue us per search, idle sampler:", launches(200), launches(200))
for name, fn in (("clock", lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                 ("reasons", lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))):
    for _ in range(50): s.search_files([df], impl, [c])
    ts = []
    for _ in range(40):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e6)
    ctx.synchronize()
    print(name, "us: median", statistics.median(ts), "max", max(ts))
import threading
stop = threading.Event()
def loop(with_reasons):
    while not stop.is_set():
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        if with_reasons: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        time.sleep(0.005)
for wr in (False, True):
    stop.clear(); t = threading.Thread(target=loop, args=(wr,), daemon=True); t.start()
    print("enqueue us per search, sampler thread with_reasons=%s:" % wr, launches(400), launches(400))
    stop.set(); t.join()
