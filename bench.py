#!/usr/bin/env python
"""Headline benchmark: batched concentrated-log-likelihood evaluations per second (BASELINE.json
configs[1]: 2-D Kriging, n = 64 Halton samples, 10^6 (theta, p) candidates = generate_grid(func, 2, 40, 25))
plus, as a secondary figure, Kriging predictions (value + MSE) per second on the kriging_r3 model.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one batch: 10^6 candidates per GPU.  N > 1 (torchrun): weak
scaling -- rank r evaluates candidates [r 10^6, (r+1) 10^6) of a finer p-grid and the only collective is
the (nll, index) all_gather of the first-minimum reduction.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'surrogate-models_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

# keep stdout to the single JSON line: NCCL prints its version banner to file descriptor 1 from C (seen on the 2-GPU
# box even with NCCL_DEBUG_FILE set), so fd 1 is pointed at stderr for the whole run and the JSON line is written to
# the saved descriptor at the end
os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + '\n').encode())

N_SAMPLES, K_DIM, T_SEC, CAND_PER_GPU = 64, 2, 40, 1000000
METRIC = 'kriging likelihood evaluations/s (n=64, k=2)'


def p_sections_for(world):
    """smallest P with 40^2 P^2 >= world * 10^6 (P = 25 at world 1: exactly the BASELINE grid)"""
    P = 25
    while (T_SEC ** K_DIM) * (P ** K_DIM) < world * CAND_PER_GPU:
        P += 1
    return P


def flops_like(n, k):
    """algorithmic flops of one likelihood evaluation, Cholesky form: n^3/3 + (k+2) n^2"""
    return n ** 3 / 3.0 + (k + 2) * n ** 2


def measured_traffic():
    """DRAM bytes of the dominant kernel per launch, from one `ncu --set full` capture of the same launch
    (profiles/r02_nll_regs_traffic.json); None if the record is absent"""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_nll_regs_traffic.json')) as f:
            return json.load(f)['traffic_bytes_per_launch']
    except Exception:
        return None


def fp64_peaks():
    """measured fp64 peaks of this pool's B200 (tools/fp64_peaks.cu -> profiles/r01_fp64_peaks.json);
    MEASURED_PEAKS.json carries HBM and bf16 only"""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r01_fp64_peaks.json')) as f:
            d = json.load(f)
        return d['dmma_m8n8k4_acc8_tflops'], d['exp_gops'], 'measured (profiles/r01_fp64_peaks.json)'
    except Exception:
        return 37.0, 818.0, 'fallback (148 SM x 64 FMA/clk x 1.965 GHz)'


# ----------------------------------------------------------------------------------------------- CPU arm
def reference_root():
    """directory of the UNMODIFIED reference when one is reachable (SM_REFERENCE=/path/to/surrogate-models, or an install
    under baseline/_ref); None on the GPU box, where the CPU arm is the oracle's statement-by-statement port"""
    for cand in (os.environ.get('SM_REFERENCE'), os.path.join(ROOT, 'baseline', '_ref')):
        if cand and os.path.exists(os.path.join(cand, 'mylibs', 'kriging.py')):
            return cand
    return None


def _ref_chunk(args):
    """the reference itself: Kriging._calc_likelihood_opti_exp per candidate (mylibs/kriging.py:130-141)"""
    lo, count, P, refdir = args
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    from oracle import numpy_oracle as orc
    from sm_b200 import workloads
    X, y = workloads.c2_halton64()
    for m in [m for m in sys.modules if m == 'mylibs' or m.startswith('mylibs.') or m == 'myutils' or m.startswith('myutils.')]:
        del sys.modules[m]
    # the reference's mylibs / myutils have no __init__.py (namespace packages): a regular package of the same name anywhere
    # on sys.path would shadow them, so this worker drops the B200 package from its path
    sys.path[:] = [refdir] + [q for q in sys.path if os.path.realpath(q or '.') != os.path.realpath(os.path.join(ROOT, 'surrogate-models_b200'))]
    from mylibs.kriging import Kriging as RefKriging
    assert os.path.realpath(sys.modules['mylibs.kriging'].__file__).startswith(os.path.realpath(refdir))
    k = RefKriging(X, y)
    k.records = None
    e, q = orc.grid_candidates(K_DIM, T_SEC, P, lo, lo + count)
    t0 = time.perf_counter()
    out = 0.0
    for c in range(count):
        v = k._calc_likelihood_opti_exp(np.concatenate([e[c], q[c]]))
        out += 0.0 if not np.isfinite(v) else v
    return time.perf_counter() - t0, out


def _cpu_chunk(args):
    """reference algorithm on `count` candidates starting at flat index lo (statement-by-statement port:
    interpreted correlation loop + np.linalg.inv + slogdet, mylibs/kriging.py:66-116)"""
    lo, count, P, mode = args
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    from oracle import numpy_oracle as orc
    from sm_b200 import workloads
    X, y = workloads.c2_halton64()
    e, q = orc.grid_candidates(K_DIM, T_SEC, P, lo, lo + count)
    f = orc.nll_lu_loops if mode == 'loops' else orc.nll_lu
    t0 = time.perf_counter()
    out = 0.0
    for c in range(count):
        try:
            v = f(X, y, [10. ** t for t in e[c]], q[c])
        except np.linalg.LinAlgError:
            v = float('inf')
        out += 0.0 if not np.isfinite(v) else v
    return time.perf_counter() - t0, out


def cpu_rate(per_proc, mode, procs, P=25, offset=250000):
    """aggregate evaluations/s of the CPU port on `procs` processes, `per_proc` candidates each, taken from
    the middle of the grid (well-conditioned and ill-conditioned candidates cost the same)"""
    import multiprocessing as mp
    fn, tag = _cpu_chunk, mode
    if mode == 'reference':
        fn, tag = _ref_chunk, reference_root()
    jobs = [(offset + i * per_proc, per_proc, P, tag) for i in range(procs)]
    t0 = time.perf_counter()
    with mp.get_context('spawn').Pool(procs) as pool:
        pool.map(fn, [(offset, 2, P, tag)] * procs)  # warm the workers (imports, BLAS init)
        t0 = time.perf_counter()
        pool.map(fn, jobs)
        dt = time.perf_counter() - t0
    return procs * per_proc / dt, dt


def _cpu_predict_chunk(args):
    """port of Kriging.predict (kriging.py:206-220) / RBF.predict (rbf.py:85-103), one interpreted call per query"""
    what, count, seed = args
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    from oracle import numpy_oracle as orc
    from sm_b200 import workloads
    rng = np.random.default_rng(seed)
    if what == 'krig':
        X, y = workloads.c5_kriging_r3()
        th, p = [0.05, 0.05], [1.9, 1.9]
        R = orc.corr_matrix_loops(X, th, p)
        Rinv = np.linalg.inv(R)
        one = np.ones(X.shape[0])
        mu = (one @ Rinv @ y) / (one @ Rinv @ one)
        xq = rng.uniform(-2, 12, (count, 2))
        t0 = time.perf_counter()
        acc = 0.0
        for x in xq:
            acc += orc.predict_lu_loops(X, y, th, p, Rinv, mu, x)
    else:
        X, y = workloads.c4_rbf4096()
        coeff = np.sin(np.arange(X.shape[0]))  # timing only: the weights do not change the cost
        xq = rng.uniform(0, 10, (count, 3))
        t0 = time.perf_counter()
        acc = 0.0
        for x in xq:
            acc += orc.rbf_predict_loops(X, coeff, 1.0, x)
    return time.perf_counter() - t0, acc


def cpu_predict_rate(what, per_proc, procs):
    import multiprocessing as mp
    with mp.get_context('spawn').Pool(procs) as pool:
        pool.map(_cpu_predict_chunk, [(what, 1, 0)] * procs)
        t0 = time.perf_counter()
        pool.map(_cpu_predict_chunk, [(what, per_proc, i + 1) for i in range(procs)])
        dt = time.perf_counter() - t0
    return procs * per_proc / dt, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = os.cpu_count()
    per_proc = 120  # ~0.6 s per process per step at ~5 ms / evaluation
    ref = reference_root()
    mode = 'reference' if ref else 'loops'
    for _ in range(args.warmup):
        cpu_rate(8, mode, cores)
    rates, t_all = [], 0.0
    for _ in range(args.steps):
        r, dt = cpu_rate(per_proc, mode, cores)
        rates.append(r)
        t_all += dt
    value = float(np.mean(rates))
    vec, _ = cpu_rate(600, 'vectorised', cores)
    sample = (f'{per_proc} candidates x {cores} processes per step from the middle of the 40x40x25x25 grid; '
              + (f'the unmodified reference at {ref} (Kriging._calc_likelihood_opti_exp)' if ref else
                 'statement-by-statement port of mylibs/kriging.py:66-116 (interpreted R loop + inv + slogdet)'))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'evals/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t_all / max(1, args.steps),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'C2: n=64 Halton samples, k=2, generate_grid(40,25) candidates (bounded sample)'},
        'cpu_baseline': {'value': value, 'unit': 'evals/s', 'cores': cores, 'kind': 'reference' if ref else 'port', 'sample': sample,
                         'vectorised_numpy_port_value': vec},
        'e2e': {'value': value, 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc = index, None
        self.path = tempfile.mktemp(prefix='clocks_', suffix='.csv')

    def start(self):
        try:
            self.fh = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                                          '-lms', '50', '-i', str(self.index)], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), sm_mhz_min=float(min(sm)))
        return out


def run_gpu(args):
    import torch
    import torch.distributed as td
    from sm_b200 import _cabi, device, dist, workloads
    from mylibs.kriging import Kriging, grid_axes

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group('nccl', device_id=torch.device('cuda', local))
    L = _cabi.lib()

    X, y = workloads.c2_halton64()
    P = p_sections_for(world)
    lo = rank * CAND_PER_GPU
    krig = Kriging(X, y)
    d = krig._dev()
    th_axis, p_axis = grid_axes(T_SEC, P)
    th_lin_h = np.array([10. ** e for e in th_axis])
    th_lin, pa = device.to_dev(th_lin_h), device.to_dev(p_axis)
    nll = device.empty(CAND_PER_GPU)
    bv, bi = device.empty(1), device.empty(1, dtype=torch.int64)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')  # 256 MiB > 126 MB L2
    stream = device.stream()

    def step():
        _cabi.check(L.sm_krig_nll_grid(device.ptr(d['X']), device.ptr(d['y']), N_SAMPLES, K_DIM, device.ptr(th_lin),
                                       T_SEC, device.ptr(pa), P, lo, CAND_PER_GPU, device.ptr(nll), device.ptr(bv),
                                       device.ptr(bi), stream))
        if world > 1:  # one 16-byte all_gather straight from the kernel's (value, index) buffers + one 16-byte read-back
            return dist.global_first_min_dev(bv, bi)
        return None

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    n0 = _cabi.launch_count()
    barrier()
    for e0, e1 in ev:
        flush.zero_()  # evict L2 between timed iterations (untimed)
        e0.record()
        step()
        e1.record()
    barrier()
    launches = _cabi.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    ms = float(np.sum([a.elapsed_time(b) for a, b in ev]))
    per_rank_ms = [ms / args.steps]
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        allms = torch.empty(world, dtype=torch.float64, device='cuda')
        td.all_gather_into_tensor(allms, t)
        per_rank_ms = [float(v) / args.steps for v in allms.tolist()]
        ms = float(allms.max().item())
    ms_per_step = ms / args.steps
    value = world * CAND_PER_GPU / (ms_per_step * 1e-3)

    # ---- kernel-only timing of the dominant kernel for the roofline (same stream, CUDA events)
    best_val, best_idx = float(bv.item()), int(bi.item())
    kms = []
    for _ in range(max(3, min(args.steps, 10))):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _cabi.check(L.sm_krig_nll_grid(device.ptr(d['X']), device.ptr(d['y']), N_SAMPLES, K_DIM, device.ptr(th_lin),
                                       T_SEC, device.ptr(pa), P, lo, CAND_PER_GPU, device.ptr(nll), None, None, stream))
        b.record()
        torch.cuda.synchronize()
        kms.append(a.elapsed_time(b))
    kernel_ms = float(np.mean(kms))  # table build (~20 us) + likelihood kernel
    peak_tf, exp_gops, peak_src = fp64_peaks()
    fl = flops_like(N_SAMPLES, K_DIM) * CAND_PER_GPU
    achieved = fl / (kernel_ms * 1e-3) / 1e12
    roofline = {
        'bound': 'tensor', 'bound_detail': 'fp64 pipe (DMMA / DFMA share one 37 TFLOP/s peak); not HBM',
        'kernel': 'krig_nll_regs_kernel<8,2,2,4,3>', 'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s',
        'frac': achieved / peak_tf, 'peak_source': peak_src, 'traffic': measured_traffic() if world == 1 else None,
        'traffic_unit': 'bytes of DRAM traffic per launch (ncu); algorithmic bytes = 8 per candidate + tables once',
        'flops_per_eval': flops_like(N_SAMPLES, K_DIM),
        'exp_per_eval': 0, 'exp_note': 'grid mode assembles R from per-dimension exponential tables: no exp per candidate, so '
                                       'the FMA-only fraction is the whole bound (no two-resource credit)',
        'kernel_ms': kernel_ms, 'hbm_gbs_achieved': (8.0 * CAND_PER_GPU) / (kernel_ms * 1e-3) / 1e9,
    }

    # ---- end to end through the host-buffer C-ABI (pinned host memory, copies inside the timed region)
    Xh = torch.from_numpy(np.ascontiguousarray(X)).pin_memory()
    yh = torch.from_numpy(np.ascontiguousarray(y)).pin_memory()
    thh = torch.from_numpy(th_lin_h).pin_memory()
    pah = torch.from_numpy(np.ascontiguousarray(p_axis)).pin_memory()
    nllh = torch.empty(CAND_PER_GPU, dtype=torch.float64).pin_memory()
    bvh = torch.empty(1, dtype=torch.float64).pin_memory()
    bih = torch.empty(1, dtype=torch.int64).pin_memory()

    def e2e_step():
        _cabi.check(L.sm_krig_nll_grid_host(Xh.data_ptr(), yh.data_ptr(), N_SAMPLES, K_DIM, thh.data_ptr(), T_SEC,
                                            pah.data_ptr(), P, lo, CAND_PER_GPU, nllh.data_ptr(), bvh.data_ptr(),
                                            bih.data_ptr()))
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
        td.all_reduce(t, op=td.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {'value': world * CAND_PER_GPU * args.steps / e2e_s, 'unit': 'evals/s',
           'h2d_bytes_per_step': int(8 * (X.size + y.size + T_SEC + P)), 'd2h_bytes_per_step': int(8 * CAND_PER_GPU + 16),
           'api': 'sm_krig_nll_grid_host (pinned host buffers)'}
    assert int(bih.item()) == best_idx

    # ---- secondary figures (device-timed, inputs resident in HBM): the other BASELINE configs at bounded size
    def timed_ms(fn, reps=2):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    # C5: predictions/s (value + MSE) on the kriging_r3 model -- the second half of BASELINE's metric
    P_FMA, P_EXP, P_POW = peak_tf * 1e12, exp_gops * 1e9, 132.26e9  # measured (profiles/r01_fp64_peaks.json)

    def wall_max(fn, reps=1):
        """wall-clock seconds of fn() bracketed by barriers, max over ranks (host-API timings)"""
        best = 1e30
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device='cuda')
                td.all_reduce(t, op=td.ReduceOp.MAX)
                dt = float(t.item())
            best = min(best, dt)
        return best

    X5, y5 = workloads.c5_kriging_r3()
    th5, p5 = np.array([0.05, 0.05]), np.array([1.9, 1.9])
    k5 = Kriging(X5, y5)
    k5.update_param(th5, p5)
    n5 = X5.shape[0]
    m = 3162  # 3162^2 ~ 1e7 mesh points per GPU over [-2, 12]^2, generated on the device
    nq5 = m * m
    xq = device.mesh_points([-2., -2.], [12., 12.], [m, m])
    pred_ms = timed_ms(lambda: k5.predict_device(xq, True))
    pred_mean_ms = timed_ms(lambda: k5.predict_device(xq, False))
    # the same mesh through the mesh entry point (per-dimension correlation factors tabulated over the axes, no coordinates read)
    mesh_ms = timed_ms(lambda: k5.predict_mesh([-2., -2.], [12., 12.], [m, m], as_device=True))
    mesh_mean_ms = timed_ms(lambda: k5.predict_mesh([-2., -2.], [12., 12.], [m, m], return_mse=False, as_device=True))
    f_mse = 2 * n5 * 2 + 2 * n5 + n5 * n5 + 3 * n5          # 2nk + 2n + n^2 + 3n flop per query with MSE
    t_pipe = nq5 * (f_mse / P_FMA + n5 / P_EXP + 2 * n5 / P_POW)   # + n exp + n k pow on the same fp64 pipe
    f_mesh = n5 * n5 + 3 * n5 + n5                            # MSE contraction + reductions + one multiply per (query, sample)
    secondary = {'metric': 'kriging predictions/s (value+MSE, n=36, k=2, p=1.9)', 'value': world * nq5 / (pred_ms * 1e-3),
                 'unit': 'predictions/s', 'queries_per_gpu': nq5, 'ms': pred_ms,
                 'value_note': 'arbitrary query points (sm_krig_predict); mesh_path = the same mesh through sm_krig_predict_mesh',
                 'mesh_path': {'value': world * nq5 / (mesh_ms * 1e-3), 'unit': 'predictions/s', 'ms': mesh_ms,
                               'kernel': 'krig_predict_mse_kernel<5,3,2,1> (mesh)', 'flops_per_query': f_mesh,
                               'frac': f_mesh * nq5 / (mesh_ms * 1e-3) / P_FMA,
                               'hbm_gbs_achieved': nq5 * 16 / (mesh_ms * 1e-3) / 1e9,
                               'mean_only_predictions_per_s': world * nq5 / (mesh_mean_ms * 1e-3),
                               'mean_only_hbm_gbs_achieved': nq5 * 8 / (mesh_mean_ms * 1e-3) / 1e9},
                 'mean_only_predictions_per_s': world * nq5 / (pred_mean_ms * 1e-3),
                 'roofline': {'bound': 'tensor', 'bound_detail': 'fp64 pipe: n^2 DMMA contraction + n exp + n k pow per query; not HBM '
                                                                 '(32 B per query)', 'kernel': 'krig_predict_mse_kernel<6,2,2,0>',
                              'flops_per_query': f_mse, 'exp_per_query': n5, 'pow_per_query': 2 * n5,
                              'achieved': f_mse * nq5 / (pred_ms * 1e-3) / 1e12, 'peak': peak_tf, 'unit': 'TFLOP/s',
                              'frac': f_mse * nq5 / (pred_ms * 1e-3) / P_FMA, 'frac_pipe_time_bound': t_pipe / (pred_ms * 1e-3),
                              'hbm_gbs_achieved': nq5 * 32 / (pred_ms * 1e-3) / 1e9, 'hbm_peak_gbs': 6532.2,
                              'traffic': None}}
    del xq
    # end to end through the host C-ABI: (a) queries generated on the device, results to pinned host memory;
    # (b) queries in pinned host memory (H2D inside the timed region), chunked over three streams
    yh5 = torch.empty(nq5, dtype=torch.float64).pin_memory()
    mh5 = torch.empty(nq5, dtype=torch.float64).pin_memory()
    X5c, y5c = np.ascontiguousarray(X5), np.ascontiguousarray(y5)
    lo5, hi5, mm5 = np.array([-2., -2.]), np.array([12., 12.]), np.array([m, m], dtype=np.int64)
    sc5 = np.empty(4)

    def e2e_mesh():
        _cabi.check(L.sm_krig_fit_predict_mesh_host(X5c.ctypes.data, y5c.ctypes.data, n5, 2, th5.ctypes.data, p5.ctypes.data,
                                                    lo5.ctypes.data, hi5.ctypes.data, mm5.ctypes.data, 0, nq5, yh5.data_ptr(),
                                                    mh5.data_ptr(), sc5.ctypes.data))
    e2e_mesh()
    s_mesh = wall_max(e2e_mesh, 2)
    xqh = torch.from_numpy(workloads.c5_query_mesh(m)).pin_memory()

    def e2e_host():
        _cabi.check(L.sm_krig_fit_predict_host(X5c.ctypes.data, y5c.ctypes.data, n5, 2, th5.ctypes.data, p5.ctypes.data,
                                               xqh.data_ptr(), nq5, yh5.data_ptr(), mh5.data_ptr(), sc5.ctypes.data))
    e2e_host()
    s_host = wall_max(e2e_host, 2)
    del xqh
    secondary['e2e'] = {'value': world * nq5 / s_mesh, 'unit': 'predictions/s', 'h2d_bytes_per_step': int(8 * (X5.size + y5.size + 4)),
                        'd2h_bytes_per_step': int(16 * nq5), 'api': 'sm_krig_fit_predict_mesh_host (fit + tabulated mesh prediction, '
                        'results to pinned host memory, 3 streams)', 'ratio_to_device_rate': (nq5 / s_mesh) / (nq5 / (mesh_ms * 1e-3)),
                        'd2h_gbs_achieved': 16 * nq5 / s_mesh / 1e9, 'note': 'bound by the device-to-host copy of the results (16 B per query)',
                        'host_queries': {'value': world * nq5 / s_host, 'h2d_bytes_per_step': int(16 * nq5), 'd2h_bytes_per_step': int(16 * nq5),
                                         'api': 'sm_krig_fit_predict_host (pinned host queries, chunked over 3 streams)'}}
    # strong scaling of north-star config 5: 1e8 queries IN TOTAL (10^4 x 10^4 mesh over [-2,12]^2) split over the ranks
    M5 = 10000
    q_lo, q_hi = dist.shard_range(M5 * M5, rank, world)

    def c5_total():
        k5.predict_mesh([-2., -2.], [12., 12.], [M5, M5], q_lo, q_hi - q_lo, as_device=True)
    c5_total()
    torch.cuda.synchronize()
    s_c5 = wall_max(c5_total, 2)
    secondary['strong_scaling_c5'] = {'queries_total': M5 * M5, 'n_gpus': world, 'seconds': s_c5, 'value': M5 * M5 / s_c5,
                                      'unit': 'predictions/s', 'what': 'Kriging.predict_mesh (value + MSE) on this rank\'s '
                                      'contiguous slice, results stay sharded on the devices; max over ranks'}
    # C3 shape: 8-D Kriging on LHS samples, explicit candidate list (list mode), n = 1024 plus two neighbours of the
    # large-n kernel range.  fractions: FMA-only (n^3/3 + (k+2) n^2 flop at the measured fp64 peak) and the
    # two-resource bound of SURVEY.md 8(d) with (k+1) n(n-1)/2 exponentials at the measured exp() rate.
    def large_n_point(n3, C3):
        X3, y3 = workloads.c3_lhs1024(n=n3)
        e3, p3 = workloads.c3_candidates(C3 * world, 8)
        k3 = Kriging(X3, y3)
        d3 = k3._dev()
        th3 = device.to_dev(np.power(10., e3[rank * C3:(rank + 1) * C3]))
        pp3 = device.to_dev(p3[rank * C3:(rank + 1) * C3])
        out3 = device.empty(C3)
        ms3 = timed_ms(lambda: _cabi.check(L.sm_krig_nll_batch(device.ptr(d3['X']), device.ptr(d3['y']), n3, 8, device.ptr(th3),
                                                                 device.ptr(pp3), C3, 0, device.ptr(out3), None, None, stream)))
        tf = flops_like(n3, 8) * C3 / (ms3 * 1e-3) / 1e12
        t_bound3 = C3 * (flops_like(n3, 8) / (peak_tf * 1e12) + 9 * n3 * (n3 - 1) / 2 / (exp_gops * 1e9))
        return {'n': n3, 'k': 8, 'value': world * C3 / (ms3 * 1e-3), 'unit': 'evals/s', 'candidates_per_gpu': C3, 'ms': ms3,
                'tflops_fma_only': tf, 'frac_of_fp64_peak': tf / peak_tf, 'frac_two_resource': t_bound3 / (ms3 * 1e-3)}
    def large_n_grid_point(n3, k3, T3, P3):
        """grid mode at large n: the candidates share |dx|^p tables, so no per-candidate powers -- the Cholesky part alone"""
        X3, y3 = workloads.c3_lhs1024(n=n3, k=k3)
        kg = Kriging(X3 * 10, y3)
        C3 = (T3 ** k3) * (P3 ** k3)
        ms3 = timed_ms(lambda: kg.likelihood_grid(T3, P3, return_nll=False))
        _, _, nllg = kg.likelihood_grid(T3, P3, return_nll=True)
        tf = flops_like(n3, k3) * C3 / (ms3 * 1e-3) / 1e12
        return {'n': n3, 'k': k3, 'mode': f'grid {T3}^{k3} x {P3}^{k3}', 'value': world * C3 / (ms3 * 1e-3), 'unit': 'evals/s',
                'candidates_per_gpu': C3, 'ms': ms3, 'tflops_fma_only': tf, 'frac_of_fp64_peak': tf / peak_tf,
                'finite_fraction': float(np.isfinite(nllg.cpu().numpy()).mean())}
    c3pt = large_n_point(1024, 1184)
    secondary_c3 = dict(c3pt, metric='kriging likelihood evaluations/s (n=1024, k=8, list mode)',
                        kernel='krig_nll_flow_kernel<6,3,1,2> (dataflow kernel, every n > 144; <4,2,1,4> below n = 1024)',
                        other_sizes=[large_n_point(256, 9472), large_n_point(512, 2368), large_n_point(2048, 296)],
                        grid_mode=large_n_grid_point(1024, 3, 8, 3),
                        # n = 128, k = 2: the CTA-per-candidate kernel with the factor in shared memory (krig_nll_cta.cu)
                        mid_n_grid_mode=large_n_grid_point(128, 2, 20, 10))
    # strong scaling of north-star config 3: 1e5 candidates IN TOTAL split over the ranks through the product API
    # (Kriging.likelihood_batch(distributed=...): H2D of this rank's rows, kernel, 16-byte argmin all_gather; results stay sharded)
    X3s, y3s = workloads.c3_lhs1024()
    e3s, p3s = workloads.c3_candidates(100000, 8)
    k3s = Kriging(X3s, y3s)
    k3s.likelihood_batch(e3s[:64 * world], p3s[:64 * world], return_best=True, gather=False)
    c3_res = {}

    def c3_total():
        _, c3_res['bv'], c3_res['bi'] = k3s.likelihood_batch(e3s, p3s, return_best=True, gather=False)
    s_c3 = wall_max(c3_total, 1)
    secondary_c3['strong_scaling_c3'] = {'candidates_total': 100000, 'n_gpus': world, 'seconds': s_c3, 'value': 100000 / s_c3,
                                         'unit': 'evals/s', 'argmin': int(c3_res['bi']), 'best_nll': float(c3_res['bv']),
                                         'what': 'Kriging.likelihood_batch(distributed) end to end: host candidates -> sharded '
                                                 'kernels -> global first minimum; wall clock, max over ranks'}
    # C4: gaussian RBF, n = 4096: weight solve + evaluation at 2e6 points per GPU
    from mylibs.rbf import RBF
    X4, y4 = workloads.c4_rbf4096()
    r4 = RBF(X4, y4)
    t0 = time.perf_counter()
    r4.update_param(1.0, 'gaus')
    torch.cuda.synchronize()
    fit_s = time.perf_counter() - t0
    r4m = RBF(X4, y4)                                            # non-definite kernel: blocked LU with cooperative panel factorisation
    r4m.update_param(0.5, 'mq')
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r4m.update_param(0.5, 'mq')
    torch.cuda.synchronize()
    fit_mq_s = time.perf_counter() - t0
    nq4 = 2000000
    xq4 = device.to_dev(workloads.c4_queries(nq4, seed=1 + rank))
    c4_ms = timed_ms(lambda: r4.predict_device(xq4))
    f_rbf = 4096 * (3 * 3 + 2)                                   # n (3k + 2) flop + n exp per query
    secondary_c4 = {'metric': 'rbf evaluations/s (gaus, n=4096, k=3)', 'value': world * nq4 / (c4_ms * 1e-3), 'unit': 'predictions/s',
                    'queries_per_gpu': nq4, 'ms': c4_ms, 'weight_solve_s': fit_s, 'weight_solve_mq_s': fit_mq_s,
                    'kernel_evals_per_s': world * nq4 * 4096 / (c4_ms * 1e-3),
                    'roofline': {'bound': 'tensor', 'bound_detail': 'fp64 pipe: 11 flop + 1 exp per (query, sample) pair; not HBM',
                                 'kernel': 'rbf_predict_kernel<3>', 'flops_per_query': f_rbf, 'exp_per_query': 4096,
                                 'achieved': f_rbf * nq4 / (c4_ms * 1e-3) / 1e12, 'peak': peak_tf, 'unit': 'TFLOP/s',
                                 'frac': f_rbf * nq4 / (c4_ms * 1e-3) / P_FMA,
                                 'frac_pipe_time_bound': nq4 * (f_rbf / P_FMA + 4096 / P_EXP) / (c4_ms * 1e-3),
                                 'hbm_gbs_achieved': nq4 * 32 / (c4_ms * 1e-3) / 1e9, 'traffic': None}}
    del xq4
    xq4h = torch.from_numpy(workloads.c4_queries(nq4, seed=1 + rank)).pin_memory()
    y4h = torch.empty(nq4, dtype=torch.float64).pin_memory()
    X4c, c4c = np.ascontiguousarray(X4), np.ascontiguousarray(r4.get_coeff())

    def e2e_rbf():
        _cabi.check(L.sm_rbf_predict_host(X4c.ctypes.data, 4096, 3, 2, 1.0, c4c.ctypes.data, xq4h.data_ptr(), nq4, y4h.data_ptr()))
    e2e_rbf()
    s_rbf = wall_max(e2e_rbf, 2)
    secondary_c4['e2e'] = {'value': world * nq4 / s_rbf, 'unit': 'predictions/s', 'h2d_bytes_per_step': int(8 * (X4.size + 4096) + 24 * nq4),
                           'd2h_bytes_per_step': int(8 * nq4), 'api': 'sm_rbf_predict_host (pinned host buffers)'}
    del xq4h

    # ---- parity record of the headline run (rank 0, BASELINE grid only): reference argmin / masked set from the committed fixtures
    parity = {'best_idx': best_idx, 'best_nll': best_val, 'gpu_argmin': best_idx}
    if rank == 0 and world == 1:
        try:
            _cabi.rescue_stats(reset=True)
            _cabi.check(L.sm_krig_nll_grid(device.ptr(d['X']), device.ptr(d['y']), N_SAMPLES, K_DIM, device.ptr(th_lin), T_SEC,
                                           device.ptr(pa), P, lo, CAND_PER_GPU, device.ptr(nll), None, None, stream))
            listed, finite = _cabi.rescue_stats()
            gnll = nll.cpu().numpy()
            gfull = np.load(os.path.join(ROOT, 'tests', 'golden', 'c2_full_f3d.npz'))
            ic = np.load(os.path.join(ROOT, 'tests', 'golden', 'c2_illcond_f3d.npz'))
            lost = np.isfinite(ic['ref']) & ~np.isfinite(gnll[ic['idx']])
            parity.update(ref_argmin=int(gfull['argmin']), index_parity=bool(int(gfull['argmin']) == best_idx),
                          n_double_double=listed, n_double_double_finite=finite,
                          n_potrf_finite_gpu_inf=int((np.isfinite(ic['chol']) & ~np.isfinite(gnll[ic['idx']])).sum()),
                          n_ref_finite_gpu_inf=int(lost.sum()),
                          n_ref_finite_gpu_inf_undercutting=int((lost & (ic['ref'] < best_val)).sum()),
                          undercutting_true_values_above_optimum=bool(np.all(ic['truth'][lost & (ic['ref'] < best_val)] > best_val)),
                          note='ref_* from tests/golden (the unmodified reference run in full); GPU-rejected candidates are those '
                               'whose fp64 correlation matrix is not positive definite in double-double; the ones the reference '
                               'scores below the optimum are LU-noise artefacts (100-digit truth above the optimum)')
        except Exception as ex:  # fixtures absent: the throughput numbers stand on their own
            parity['note'] = f'parity fixtures unavailable: {ex}'
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count()
            per_proc = 3000
            v_loops, dt1 = cpu_rate(per_proc, 'loops', cores)
            v_vec, dt2 = cpu_rate(1500, 'vectorised', cores)
            vp, dtp = cpu_predict_rate('krig', 120000, cores)
            secondary['cpu_baseline'] = {'value': vp, 'unit': 'predictions/s', 'cores': cores, 'kind': 'port',
                                         'sample': f'120000 queries x {cores} processes ({dtp:.1f} s); statement-by-statement port of '
                                                   'Kriging.predict (mylibs/kriging.py:206-220; mean only -- the reference has no MSE)'}
            vr, dtr = cpu_predict_rate('rbf', 1200, cores)
            secondary_c4['cpu_baseline'] = {'value': vr, 'unit': 'predictions/s', 'cores': cores, 'kind': 'port',
                                            'sample': f'1200 queries x {cores} processes ({dtr:.1f} s); statement-by-statement port of '
                                                      'RBF.predict (mylibs/rbf.py:85-103), n = 4096'}
            cpu = {'value': v_loops, 'unit': 'evals/s', 'cores': cores, 'kind': 'port',
                   'sample': f'{per_proc} candidates x {cores} processes from the middle of the grid ({dt1:.1f} s); '
                             'statement-by-statement port of mylibs/kriging.py:66-116 (interpreted R loop, inv, slogdet)',
                   'vectorised_numpy_port_value': v_vec,
                   'vectorised_sample': f'1500 candidates x {cores} processes ({dt2:.1f} s); numpy-vectorised R build + inv + slogdet'}
        line = {
            'metric': METRIC, 'value': value, 'unit': 'evals/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': f'C2: n=64 Halton samples on [0,10]^2 (y=f_3d), k=2, 1e6 (theta,p) candidates per GPU '
                                   f'= generate_grid(func,2,{T_SEC},{P}) flat range [r*1e6,(r+1)*1e6)',
                       'candidates_per_gpu': CAND_PER_GPU, 'l2': 'flushed (256 MiB write) between timed steps',
                       'parallelism': f'candidate-sharded x{world}; all_gather of (nll_min, index) only'},
            'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'per_rank_ms_per_step': per_rank_ms,
            'cpu_baseline': cpu, 'secondary': secondary, 'secondary_c3': secondary_c3, 'secondary_c4': secondary_c4,
            'result': parity,
        }
        emit(line)
    if world > 1:
        td.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    return run_gpu(args)


if __name__ == '__main__':
    sys.exit(main())
