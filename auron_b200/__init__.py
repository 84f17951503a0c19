"""auron_b200 -- B200-native execution engine behind Apache Auron's native-engine boundary.

The product is `libauron_b200.so` (hand-written sm_100a CUDA + C++ host runtime, csrc/).  This package
only holds the thin host-side binding (`runtime`) and the plan wire-format encoder (`proto`) that play
the role of Auron's JVM glue in tests and benchmarks.  Nothing here computes on the CPU.
"""
from . import proto, runtime  # noqa: F401
from .runtime import AuronError, Task, run_task  # noqa: F401
