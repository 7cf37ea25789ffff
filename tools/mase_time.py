"""Times the MASE / BASE tails alone (the bench's `mase_base` extra workload) and prints its JSON."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from active_learning_b200.engine import Engine  # noqa: E402

peak, _src = bench.measured_peaks()
print(json.dumps(bench.run_mase_workload(Engine(), peak, steps=3, warmup=1)))
