"""Dev tool: non-definite RBF weight solve (blocked LU) with the cooperative panel kernel against the single-CTA panel
(SMB_LU_COOP_OFF=1, subprocess): coefficients must agree bit for bit (same pivots, same arithmetic); wall time of both."""
import json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
if len(sys.argv) > 1 and sys.argv[1] == 'worker':
    import torch
    from mylibs.rbf import RBF
    from sm_b200 import workloads
    out = {}
    for n in (900, 1500, 2500, 4096):
        X, y = workloads.c4_rbf4096(n)
        for name, a in (('mq', 0.5), ('lin', 1.0), ('cubic', 1.0), ('imq', 0.5)):
            r = RBF(X, y); r.update_param(a, name); torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); r.update_param(a, name); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
            c = np.asarray(r.get_coeff() if hasattr(r, 'get_coeff') else r._coeff)
            out[f'{n}_{name}'] = dict(ms=min(ts) * 1e3, coeff=c.view(np.int64).tolist() if n <= 1500 else [], head=c[:64].view(np.int64).tolist(),
                                      s=float(np.sum(np.abs(c))))
    print(json.dumps(out)); sys.exit(0)
res = {}
for flag in ('0', '1'):
    r = subprocess.run([sys.executable, os.path.abspath(__file__), 'worker'], env=dict(os.environ, SMB_LU_COOP_OFF=flag), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    res[flag] = json.loads(r.stdout.strip().split('\n')[-1])
for k in res['0']:
    a, b = res['0'][k], res['1'][k]
    same = a['head'] == b['head'] and a['coeff'] == b['coeff'] and a['s'] == b['s']
    print(f'{k:12s} cooperative panel {a["ms"]:8.2f} ms | single-CTA panel {b["ms"]:8.2f} ms | coefficients bit-identical: {same}', flush=True)
    assert same
