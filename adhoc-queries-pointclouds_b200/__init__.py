"""B200-native full-scan query path of igd-geo/adhoc-queries-pointclouds.

The directory name carries the reference's name (hyphens included), so import it with
``importlib.import_module("adhoc-queries-pointclouds_b200")`` (see ``pcq_import.py`` at the repo
root).  The product is ``libpcq.so`` (csrc/: hand-written sm_100a CUDA kernels behind the C ABI of
include/pcq.h) plus the `query` CLI; this package is the Python twin of the reference's
Searcher / ResultCollector interface that the tests and bench.py drive it through.
"""
from . import binding, group, sharding, synth  # noqa: F401
from .group import Dataset, Group, Result, shard_plan  # noqa: F401
from .binding import (  # noqa: F401
    CANDIDATE_DTYPE,
    POINT_DTYPE,
    FileDesc,
    PcqError,
    Query,
    SynthSpec,
    lib,
)
from .searcher import (  # noqa: F401
    BoundsSearcher,
    BufferCollector,
    ClassSearcher,
    Context,
    CountCollector,
    DeviceFile,
    GridSampledCollector,
    HostIndex,
    ResultCollector,
    SearchImplementation,
    Searcher,
    default_context,
    index_filter,
    run_search_parallel,
    run_search_sequential,
    reset_collectors,
    search_host_files_multi,
    search_las_file_by_bounds_optimized,
    search_las_file_by_classification_optimized,
    search_last_file_by_bounds_optimized,
    search_last_file_by_classification_optimized,
)
