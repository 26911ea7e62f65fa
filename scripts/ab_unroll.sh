#!/bin/bash
# A/B of compile-time experiment knobs on one box: one build per argument (SHMFAST_NVCC_EXTRA), "" = the default build.
set -u
run() {
  python scripts/prof_tc_ol.py 2>&1 | head -1
  python scripts/prof_tc.py 2>&1 | head -1
  python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('4dof', round(d['value']), d['roofline']['kernel_ms'], d['clocks']['sm_mhz'])"
  python bench.py --workload openlab_hybrid --steps 20 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('openlab', round(d['value']), d['clocks']['sm_mhz'])"
}
for v in "$@"; do echo "== build: [$v]"; export SHMFAST_NVCC_EXTRA="$v"; run; done
