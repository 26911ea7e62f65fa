#!/bin/bash
# Same-box A/B of one compile-time knob ($1) against the default build: GPU tests on the variant, then the two hybrid benches each.
set -u
b() {
  timeout 60 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('4dof', round(d['value']), d['roofline']['kernel_ms'], d['clocks']['sm_mhz'])"
  timeout 60 python bench.py --workload openlab_hybrid --steps 20 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('openlab', round(d['value']), d['clocks']['sm_mhz'])"
}
export SHMFAST_NVCC_EXTRA="$1"
echo "== variant [$1]"; timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -2; b
export SHMFAST_NVCC_EXTRA=""
echo "== default"; b
