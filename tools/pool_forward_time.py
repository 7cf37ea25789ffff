"""Runs the bench's end-to-end pool-forward workload alone and prints its JSON."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from active_learning_b200.engine import Engine  # noqa: E402

print(json.dumps(bench.run_pool_forward_workload(Engine())))
