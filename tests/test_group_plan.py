"""CPU tests of the host side of the multi-GPU layer (include/pcq.h: pcq_shard_plan, pcq_group_*): the shard planner
is exact and total, and a group fails loudly without a GPU.  The data path of a group is covered on the GPU box by
tests/test_gpu_group.py (one process, n members) and tests/dist_group_nccl.py (torchrun, NCCL)."""
import subprocess

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

CHUNK = 8192


@settings(max_examples=200, deadline=None, derandomize=True)
@given(st.lists(st.integers(0, 400_000), min_size=0, max_size=12), st.integers(1, 9))
def test_range_plan_covers_every_point_exactly_once(pcq, ppf, world):
    plan = pcq.shard_plan(ppf, world, pcq.binding.SHARD_RANGES)
    for f, n in enumerate(ppf):
        mine = plan[plan["file"] == f]
        # ascending with the rank, contiguous, starting on chunk boundaries: the concatenation of the members' record
        # streams in rank order IS the file's scan order (BufferCollector, collect_points.rs:29-31)
        assert list(mine["rank"]) == sorted(mine["rank"]) and len(set(mine["rank"])) == len(mine)
        pos = 0
        for sh in mine:
            assert int(sh["first_point"]) == pos and int(sh["first_point"]) % CHUNK == 0 and int(sh["n_points"]) > 0
            pos += int(sh["n_points"])
        assert pos == n
    assert np.all(plan["rank"] < world)
    # balanced to within one chunk per file
    if ppf:
        load = [int(plan["n_points"][plan["rank"] == r].sum()) for r in range(world)]
        assert max(load) - min(load) <= CHUNK * len(ppf)


@settings(max_examples=200, deadline=None, derandomize=True)
@given(st.lists(st.integers(0, 400_000), min_size=0, max_size=12), st.integers(1, 9))
def test_file_plan_keeps_files_whole_and_balances(pcq, ppf, world):
    plan = pcq.shard_plan(ppf, world, pcq.binding.SHARD_FILES)
    assert sorted(plan["file"]) == [f for f, n in enumerate(ppf) if n > 0]
    for sh in plan:
        assert int(sh["first_point"]) == 0 and int(sh["n_points"]) == ppf[int(sh["file"])]
    if ppf:
        load = [int(plan["n_points"][plan["rank"] == r].sum()) for r in range(world)]
        assert max(load) - min(load) <= max(ppf)  # greedy largest-first bound


def test_doc_s_box_spreads_over_all_gpus(pcq):
    """the doc-S box touches 5 of 64 tiles; with range sharding every GPU still holds an eighth of each."""
    plan = pcq.shard_plan([31_250_000] * 64, 8)
    hit = [8, 16, 24, 32, 40]
    per_rank = [int(plan["n_points"][(plan["rank"] == r) & np.isin(plan["file"], hit)].sum()) for r in range(8)]
    assert sum(per_rank) == 5 * 31_250_000 and max(per_rank) - min(per_rank) <= 5 * CHUNK


def test_plan_rejects_bad_arguments(pcq):
    with pytest.raises(pcq.PcqError):
        pcq.shard_plan([10], 0)
    with pytest.raises(pcq.PcqError):
        pcq.shard_plan([10], 2, mode=7)


def test_group_needs_a_gpu_and_says_so(pcq):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pcq.PcqError) as e:
        pcq.Group.local(2)
    assert e.value.code == pcq.binding.PCQ_ERR_CUDA and "no CPU fallback" in e.value.message


def test_library_does_not_link_nccl(pcq):
    """NCCL is resolved at run time (dlopen), so that libpcq.so loads on a box without it and shares the copy a host
    framework has already loaded."""
    out = subprocess.check_output(["readelf", "-d", str(pcq.binding.LIB_PATH)], text=True)
    assert "nccl" not in out.lower()
