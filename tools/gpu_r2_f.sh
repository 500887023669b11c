#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampler.py -x -q -k "deterministic or quickly" 2>&1 | tail -15
timeout 600 python - <<'PY' 2>&1 | tail -30
import time, numpy, torch, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from xenoverse_b200.task_sampler import sample_mdp_candidates, check_valuefunction_device
for ns, count in ((16, 2048), (32, 2048), (64, 1024), (128, 1024)):
    torch.cuda.synchronize(); t0 = time.time()
    c = sample_mdp_candidates(count, ns, 5, seed=1)
    torch.cuda.synchronize(); t1 = time.time()
    ms = numpy.random.default_rng(0).uniform(max(4*ns,100), max(min(8*ns,500), max(4*ns,100)+1), count)
    r = check_valuefunction_device(c, ns, ms)
    torch.cuda.synchronize(); t2 = time.time()
    ok = (c['status'] == 0)
    print(f"ns={ns}: {count} candidates: gen {t1-t0:.2f} s (sweeps mean {c['sweeps'].float().mean():.0f}), filter {t2-t1:.2f} s (sweeps mean {r['sweeps'].float().mean():.0f}); ok {ok.float().mean():.3f} accept|ok {r['accept'][ok].float().mean():.3f}"
          f" gap {r['min_gap'][ok].median():.2f} gini {r['gini'][ok].median():.3f} ent {r['entropy'][ok].median():.3f} launches {r['launches']}")
PY
timeout 600 python -m pytest tests/test_gpu_policy.py tests/test_task_sampler.py -x -q 2>&1 | tail -5
