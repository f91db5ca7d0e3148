#!/usr/bin/env python
"""bench.py -- FP64 cell-updates/s of the rhoSolver hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm

Workload (BASELINE.json configs[3], the one the metric is quoted on): synthetic
unit-cube tet mesh, 203^3 hexes x 6 Kuhn tets = 50 192 562 cells, Roe flux,
second order, explicit, all walls, smooth acoustic initial state
rho = p = 1 + 0.1 sin(2 pi x) sin(2 pi y) sin(2 pi z) (SURVEY 8d's SOD split goes
NaN under the reference's unlimited scheme on Kuhn tets; it is the
`--shock 1 --limiter bj --cfl 0.4` line), DT = 1e-4.  A "step" is one
RhoSolver::solve() + residual + new->old over the whole mesh.

Both arms print the same `config`; `--impl reference` runs the CPU restatement on
ALL host cores (whatever WORLD_SIZE says), on the same mesh at the same size.

What is timed: `value` = ONE un-instrumented mstgpu_step(dt, K) call (CUDA events on the
solver's stream, max over ranks); the per-kernel durations of `roofline` come from an
instrumented repeat of the same K steps; `e2e` = one mstgpu_step_host call per step with
pinned HOST buffers on both sides (H2D, tiles and D2H pipelined) + the residual, with the
same sequence as three separate calls beside it; `every_step_residual` = K calls of one
step.  `secondary` holds short runs of BASELINE configs 1, 2, 3 and 5 through the same code.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

# torchrun pins OMP_NUM_THREADS=1; the host-side mesh / plan / partition builders
# are OpenMP code, so give every rank its share of the cores before libgomp loads
_world = int(os.environ.get("WORLD_SIZE", "1"))
_ref_arm = any(a == "reference" or a == "--impl=reference" for a in sys.argv[1:])
if _ref_arm:
    # the reference arm runs on rank 0 alone and is the CPU baseline of every N: all host cores, always
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
elif os.environ.get("OMP_NUM_THREADS", "1") == "1":
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // _world))

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "mst-cfd_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

DT = 1e-4
# algorithmic bytes per cell-update (every distinct datum once per pass)
ALGO_BYTES = {
    (3, 2): dict(step=600, gradient=248, flux_update=352),
    (3, 1): dict(step=160, gradient=0, flux_update=160),
    (2, 2): dict(step=370, gradient=152, flux_update=218),
    (2, 1): dict(step=114, gradient=0, flux_update=114),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def sod_state(f):
    """Time::initialization (R/time/Time.cpp:13-38) on the base state of CONST.h:70-75"""
    U = f["dim"] + 2
    Q = np.zeros((f["ncells"], U))
    Q[:, 0] = 1.0
    Q[:, U - 1] = 1 * (1 / 286.32 * 715.8 + 0.5 * (0 * 0 + 0 * 0))
    right = f["cc"][:, 0] > 0.5
    Q[right, 0] *= 0.125
    Q[right, U - 1] *= 0.1
    return Q


def build_workload(args):
    """Returns (flat mesh, Q0, description)."""
    from mstgpu import host
    t = time.time()
    if args.workload == "box":
        n = args.n
        raw = host.box_tets_raw(n, n, n)
        f = host.flatten_raw(raw)
        x = f["cc"]
        Q = np.zeros((f["ncells"], 5))
        # Smooth acoustic initial state (every face carries a jump).  SURVEY.md 8d
        # proposed a SOD split + perturbation; the reference scheme has no
        # limiter and that state goes NaN within ~15 steps on Kuhn tets at any
        # DT (measured, DESIGN.md "Measurement"), so the bench uses the smooth
        # part only -- arithmetic per face is the same.
        pert = 0.1 * np.sin(2 * np.pi * x[:, 0]) * np.sin(2 * np.pi * x[:, 1]) * np.sin(2 * np.pi * x[:, 2])
        Q[:, 0] = 1.0 + pert
        Q[:, 4] = (1.0 + pert) / 0.4
        desc = (f"unit-cube {n}^3 hexes x 6 Kuhn tets, Roe, 2nd order, explicit, all walls, "
                "smooth acoustic init rho=p=1+0.1*sin*sin*sin")
        if getattr(args, "shock", 0):
            # SURVEY 8d's input: SOD split + smooth perturbation (runs with the limiter extension)
            Q[:, 0] = np.where(x[:, 0] > 0.5, 0.125, 1.0) + 0.1 * pert
            Q[:, 4] = (np.where(x[:, 0] > 0.5, 0.1, 1.0) + 0.1 * pert) / 0.4
            desc = (f"unit-cube {n}^3 hexes x 6 Kuhn tets, Roe, 2nd order + limiter, all walls, "
                    "SOD split at x=0.5 + 1e-2*sin*sin*sin on rho and p")
    elif args.workload == "step":
        raw = host.forward_step_raw(args.n)
        f = host.flatten_raw(raw)
        u = 3.0 * np.sqrt(1.4)
        Q = np.tile(np.array([1.0, u, 0.0, 1.0 / 0.4 + 0.5 * u * u]), (f["ncells"], 1))
        desc = f"forward-facing step h=1/{args.n}, triangles"
    elif args.workload == "sphere":
        # BASELINE config 3: flow over a sphere, cubed-sphere shell, 24 tets per hex; Mach 0.5 in +x
        n = args.n if args.n != 203 else 42
        m = max(2, round(n * 40 / 42))
        f = host.flatten_raw(host.sphere_shell_raw(n, m))
        u = 0.5 * np.sqrt(1.4)
        Q = np.tile(np.array([1.0, u, 0.0, 0.0, 1.0 / 0.4 + 0.5 * u * u]), (f["ncells"], 1))
        desc = (f"cubed-sphere shell r in [0.5, 10], 6 x {n}^2 x {m} hexes x 24 tets, sphere = wall, outer = inlet / outlet, "
                "Mach 0.5 free stream")
    elif args.workload == "sod":
        # BASELINE config 1: the reference's own SOD tube mesh (18 282 triangles; the raw tables of the shipped
        # file are a test fixture), written as a .msh file once and read back by the native reader, so that both
        # arms start from the same file like the reference program does
        d = np.load(os.path.join(ROOT, "tests", "golden", "mesh_2d-shockwavepipe-2.npz"))
        raw = dict(dim=int(d["dim"]), ncells=int(d["ncells"]), nodes=d["nodes"], face_nodes=d["face_nodes"], c0=d["c0"], c1=d["c1"],
                   zones=[dict(start=int(a), end=int(b), type=int(t)) for a, b, t in zip(d["zone_start"], d["zone_end"], d["zone_type"])])
        args.mesh = os.path.join(tempfile.mkdtemp(prefix="mstbench_"), "2d-shockwavepipe-2.msh")
        host.write_msh(args.mesh, raw)
        f = host.flatten_raw(host.read_msh(args.mesh))
        Q = sod_state(f)
        desc = "2d-shockwavepipe-2.msh (the reference's SOD case), Roe, 2nd order, explicit, DT = 1/STEP_TIME"
    elif args.workload == "msh":
        # any Fluent .msh file the reference's reader accepts (native reader, host/mshread.cpp), with the
        # reference's SOD initial state (Time.cpp:13-38): BASELINE config 1 is --mesh .../2d-shockwavepipe-2.msh
        if not args.mesh:
            raise SystemExit("--workload msh needs --mesh FILE")
        f = host.flatten_raw(host.read_msh(args.mesh))
        Q = sod_state(f)
        desc = f"{os.path.basename(args.mesh)} (native .msh reader), SOD initial state, DT = 1/STEP_TIME"
    else:
        raise SystemExit("unknown workload")
    log(f"[bench] mesh: {f['ncells']} cells, {f['nfaces']} faces in {time.time() - t:.1f}s")
    return f, Q, desc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (recipe line
    of B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,clocks.mem,temperature.gpu")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, pw, mem, temp = [], [], set(), [], [], []
        for ln in self.f.read().splitlines():
            c = [t.strip() for t in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            try:
                mem.append(float(c[9])); temp.append(float(c[10]))
            except (ValueError, IndexError):
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)),
                    mem_mhz=float(np.median(mem)) if mem else None, temp_c_max=float(max(temp)) if temp else None,
                    samples=len(sm), reasons=sorted(reasons))


def run_params(args, Q0):
    """(dt, inlet state) of a workload: the inlet of the step / sphere cases is their free stream"""
    dt = {"box": DT, "step": 2e-5, "sphere": 2e-4, "msh": 2.5e-4, "sod": 2.5e-4}.get(args.workload, DT)
    inlet = (list(Q0[0]) + [0.0] * 5)[:5] if args.workload in ("step", "sphere") else None
    return dt, inlet


def cpu_sample_size(args):
    return {"box": args.cpu_n, "step": max(args.cpu_n, 200), "sphere": min(args.cpu_n, 20)}.get(args.workload, args.cpu_n)


def scheme_kwargs(args):
    """gradient / limiter choice (extension; defaults = the reference scheme)"""
    return dict(gradient=args.gradient, limiter=args.limiter, limiter_k=args.limiter_k)


def workload_config(args, world, ncells, nfaces, U, desc, dt_run):
    """`config` of the JSON line: names the workload.  Built by this one function for BOTH arms, so the
    reference arm's line describes the same job (same mesh, size, scheme, dt) as the GPU arm's."""
    return dict(workload=f"{args.workload}{args.n}: {desc}", cells=int(ncells), faces=int(nfaces), flux=args.flux,
                order=args.order, viscous=args.viscous, gradient=args.gradient, limiter=args.limiter, cfl=args.cfl,
                implicit=None if not args.implicit else dict(dt=args.implicit_dt, lusgs_iterations=LUSGS_ITERS, sweeps="colour-ordered"),
                dt=dt_run,
                parallelism=f"{world} partition(s) on {world} GPU(s), Hilbert ranges, 2 ghost layers, halo exchange + allreduce(max); "
                            "the CPU arm runs the whole mesh on one host",
                l2="inputs larger than L2 (state + tables >> 126 MB)" if ncells * U * 8 > 2 ** 28
                else "inputs smaller than L2: flush not applied")


_M1, _M2, _GOLD = np.uint64(0xBF58476D1CE4E5B9), np.uint64(0x94D049BB133111EB), np.uint64(0x9E3779B97F4A7C15)


def state_digest(Q, gids):
    """Order-independent 64-bit digest of a set of state rows keyed by GLOBAL cell id: the sum mod 2^64 of
    splitmix64(bits(Q[i, k]) xor golden * (gid_i * U + k + 1)).  Ranks add their partial sums, so equal
    digests at N = 1, 2, 4, 8 mean every conserved variable of every cell has the same 64 bits."""
    U = Q.shape[1]
    tot = np.uint64(0)
    with np.errstate(over="ignore"):
        for a in range(0, Q.shape[0], 1 << 22):
            q = np.ascontiguousarray(Q[a:a + (1 << 22)]).view(np.uint64)
            key = (gids[a:a + (1 << 22)].astype(np.uint64)[:, None] * np.uint64(U) + np.arange(1, U + 1, dtype=np.uint64)[None, :]) * _GOLD
            z = q ^ key
            z = (z ^ (z >> np.uint64(30))) * _M1
            z = (z ^ (z >> np.uint64(27))) * _M2
            z = z ^ (z >> np.uint64(31))
            tot = tot + z.sum(dtype=np.uint64)
    return int(tot)


def cpu_baseline(args, nsteps=None, n=None, budget_s=12.0):
    """Oracle (kind "port": the reference cannot be compiled without Eigen) on
    the host cores, on a bounded sample of the same workload: about `budget_s`
    seconds of CPU work (the step count is sized from one warm-up step)."""
    from oracle import oracle
    from mstgpu import host
    n = n or cpu_sample_size(args)
    a = argparse.Namespace(**vars(args)); a.n = n
    f, Q, _ = build_workload(a)
    DT, inlet = run_params(args, Q)
    o = oracle.Oracle(f, order=args.order, flux=args.flux, nthreads=0, viscous=args.viscous, inletQ=inlet, **scheme_kwargs(args))
    if getattr(args, "implicit", 0):
        # implicit step: OpenMP assembly + the reference's sequential block LU-SGS (lusgs_oracle.cpp)
        def run(k, Q=Q):
            for _ in range(k):
                Q = o.step_implicit(args.implicit_dt, Q, LUSGS_ITERS)
    else:
        run = lambda k: o.run(DT, k, Q)
    run(1)  # warm-up (page faults of the work arrays)
    t = time.perf_counter()
    run(1)
    t1 = time.perf_counter() - t
    if nsteps is None:
        nsteps = int(min(60, max(1 if getattr(args, "implicit", 0) else 3, round(budget_s / max(t1, 1e-3)))))
    t = time.perf_counter()
    run(nsteps)
    el = time.perf_counter() - t
    return dict(value=f["ncells"] * nsteps / el, unit="cell-updates/s", cores=o.nthreads, kind="port",
                sample=f"{args.workload} n={n}: {f['ncells']} cells x {nsteps} steps in {el:.2f}s "
                       f"(oracle/rho_oracle.cpp, OpenMP, host has {os.cpu_count()} cpus)")


def run_reference(args, rank):
    if rank != 0:
        return
    if args.workload == "lusgs":
        c = lusgs_cpu(args, budget_s=max(5.0, 10.0 * args.steps / 20))
        out = dict(impl="reference", metric="cell_updates_per_sec", value=c["value"], unit="cell-updates/s",
                   n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=None, higher_is_better=True,
                   scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                   config=dict(workload=f"lusgs box{args.n}: block-5 LU-SGS, {LUSGS_ITERS} iterations per solve (bounded sample)"),
                   cpu_baseline=c, e2e=dict(value=c["value"], unit="cell-updates/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                   gpu_launches=0)
        emit(json.dumps(out))
        return
    ref_io = os.path.join(ROOT, "oracle", "_ref", "ref_io")
    if args.workload == "sod" and args.order == 2 and args.flux == "roe" and os.path.exists(ref_io):
        # BASELINE config 1 is the one case the reference itself can run: its own reader, Time::goNextTimeStep
        # and RhoSolver (Roe, ACCURACY 2; oracle/_ref/ref_io = the reference's sources, oracle/refbuild), with
        # the 8 OpenMP threads it hard-codes (CONST.h:7), on the same .msh file
        f, Q, desc = build_workload(args)
        tmp = os.path.dirname(args.mesh)
        os.makedirs(os.path.join(tmp, "result"), exist_ok=True)
        nst = args.steps + min(args.warmup, 3)
        r = subprocess.run([ref_io, args.mesh, tmp, "-", "0", str(nst), "1"], check=True, capture_output=True, text=True)
        tok = r.stdout.split()
        el = float(tok[tok.index("step_ms") + 1]) * 1e-3
        thr = int(tok[tok.index("threads") + 1])
        val = f["ncells"] * nst / el
        out = dict(impl="reference", metric="cell_updates_per_sec", value=val, unit="cell-updates/s",
                   n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * el / nst,
                   higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="reference mesh",
                   config=dict(workload=f"sod: " + desc, flux=args.flux, order=args.order, dt=2.5e-4),
                   cpu_baseline=dict(value=val, unit="cell-updates/s", cores=thr, kind="reference",
                                     sample=f"{f['ncells']} cells x {nst} steps of Time::goNextTimeStep in {el:.3f}s "
                                            f"(oracle/_ref/ref_io, {thr} OpenMP threads as hard-coded, host has {os.cpu_count()} cpus)"),
                   e2e=dict(value=val, unit="cell-updates/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                   gpu_launches=0)
        emit(json.dumps(out))
        return
    from oracle import oracle
    # the SAME workload at the SAME size as the GPU arm (50.2 M tets by default: ~3-5 s per step on the box's
    # cores, so --steps 20 ends within two minutes); --cpu-sample N runs an N^3 sample instead and says so
    a = argparse.Namespace(**vars(args))
    sampled = args.cpu_sample > 0
    if sampled:
        a.n = args.cpu_sample
    f, Q, desc = build_workload(a)
    DT, inlet = run_params(args, Q)
    if args.implicit:
        DT = args.implicit_dt
    o = oracle.Oracle(f, order=args.order, flux=args.flux, nthreads=0, viscous=args.viscous, inletQ=inlet, **scheme_kwargs(args))
    if args.implicit:
        def adv(k, Q):
            for _ in range(k):
                Q = o.step_implicit(args.implicit_dt, Q, LUSGS_ITERS)
            return Q
    else:
        adv = lambda k, Q: o.run(DT, k, Q, residuals=True)[0]  # Time.cpp:69-76: the residual of every step
    nw = min(args.warmup, 2)  # page faults of the work arrays; a CPU has no clocks to ramp
    Q = adv(nw, Q)
    t = time.perf_counter()
    Q = adv(args.steps, Q)
    el = time.perf_counter() - t
    val = f["ncells"] * args.steps / el
    sample = (f"{a.workload} n={a.n}: {f['ncells']} cells x {args.steps} timed steps (+{nw} warm-up) in {el:.2f}s, "
              f"{'a SAMPLE of the n=%d workload, ' % args.n if sampled else 'the full workload, '}"
              f"oracle/rho_oracle.cpp, {o.nthreads} OpenMP threads of {os.cpu_count()} cpus, residual every step")
    log(f"[bench] reference arm: {sample}")
    out = dict(impl="reference", metric="cell_updates_per_sec", value=val, unit="cell-updates/s",
               n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * el / args.steps,
               higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
               config=workload_config(a, args.gpus, f["ncells"], f["nfaces"], f["dim"] + 2, desc, DT),
               cpu_baseline=dict(value=val, unit="cell-updates/s", cores=o.nthreads, kind="port", sample=sample),
               e2e=dict(value=val, unit="cell-updates/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               gpu_launches=0)
    emit(json.dumps(out))


def lusgs_system(args, n):
    """Pattern of an implicit operator on the box mesh (device cell order) + its colour sweep order."""
    import mstgpu
    from mstgpu import host
    t = time.time()
    f = host.flatten_raw(host.box_tets_raw(n, n, n))
    rowptr, col = mstgpu.mesh_adjacency(f)
    order, ncol = mstgpu.lusgs_color_order(rowptr, col)
    rows = np.repeat(np.arange(rowptr.size - 1, dtype=np.int32), np.diff(rowptr))
    dpos = np.nonzero(col == rows)[0]
    log(f"[bench] lusgs pattern: {rowptr.size - 1} rows, {col.size} blocks, {ncol} colours in {time.time() - t:.1f}s")
    return f, rowptr, col, order, ncol, dpos


LUSGS_ITERS = 5   # LU_INTERVAL, R/include/CONST.h:58
LUSGS_B = 5       # DIMU of a 3-D mesh


def lusgs_bytes_per_row(nnz_per_row):
    """per iteration the off-diagonal blocks are read twice (val for U x / L ux, the
    D^-1-scaled copies in the sweeps), the diagonal blocks three times, ~7 vector passes; once per
    solve the blocks are read and their scaled copies / D / D^-1 written."""
    BB, off = 8 * LUSGS_B * LUSGS_B, nnz_per_row - 1.0
    per_iter = 2 * off * BB + 3 * BB + 7 * 8 * LUSGS_B
    setup = nnz_per_row * BB + off * BB + 2 * BB
    return per_iter, setup


def run_lusgs(args, rank, world):
    """BASELINE config 5's sweep: the reference's block LU-SGS (SparseSolver<MT,VCT>::solveILU,
    R/lusolver/SparseSolver.cpp:54-104) on the 50 M-tet adjacency, colour-ordered, 5 iterations,
    device-resident.  One "cell update" = one row through one whole solve."""
    import torch
    import mstgpu
    if world > 1:
        raise SystemExit("bench.py --workload lusgs: replicas only (the sweep does not shard yet, DESIGN.md 5)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(0)
    f, rowptr, col, order, ncol, dpos = lusgs_system(args, args.n)
    n, nnz, B = rowptr.size - 1, col.size, LUSGS_B
    t = time.time()
    solver = mstgpu.LuSgs(rowptr, col, B, device=0, sweep_order=order)
    lev = solver.levels()
    g = torch.Generator(device="cuda"); g.manual_seed(20231017)
    val = torch.empty((nnz, B, B), dtype=torch.float64, device="cuda")
    step = 1 << 22
    for i in range(0, nnz, step):  # diagonally dominant random blocks, generated in place on the device
        val[i:i + step] = (torch.rand((min(step, nnz - i), B, B), generator=g, dtype=torch.float64, device="cuda") - 0.5) * 0.2
    eye = torch.eye(B, dtype=torch.float64, device="cuda") * 3.0
    dp = torch.from_numpy(dpos).cuda()
    for i in range(0, n, step):
        val[dp[i:i + step]] += eye
    b = torch.rand((n, B), generator=g, dtype=torch.float64, device="cuda")
    x0 = torch.ones((n, B), dtype=torch.float64, device="cuda")
    x = x0.clone()
    torch.cuda.synchronize()
    dev_bytes = solver.device_bytes + val.numel() * 8 + 3 * n * B * 8
    log(f"[bench] lusgs solver built in {time.time() - t:.1f}s, levels fwd/bwd {lev}, {dev_bytes / 2**30:.1f} GiB on device")
    for _ in range(args.warmup):
        x.copy_(x0); solver.solve_device(val.data_ptr(), b.data_ptr(), x.data_ptr(), LUSGS_ITERS)
    clocks = ClockSampler(0); clocks.start(); time.sleep(0.3)
    l0 = solver.launch_count
    ms = 0.0
    w0 = time.perf_counter()
    for _ in range(args.steps):  # every solve starts from the same x0 (untimed reset), as every time step would
        x.copy_(x0); torch.cuda.synchronize()
        ms += solver.solve_device(val.data_ptr(), b.data_ptr(), x.data_ptr(), LUSGS_ITERS)
    wall = time.perf_counter() - w0
    clk = clocks.stop()
    launches = solver.launch_count - l0
    value = n * args.steps / (ms * 1e-3)
    # fixed point: the iteration converges to A x = b; the residual of the block system after 5 sweeps
    peak, peak_src = measured_peaks()
    per_iter, setup = lusgs_bytes_per_row(nnz / n)
    algo = LUSGS_ITERS * per_iter + setup
    achieved = algo * n / (ms / args.steps * 1e-3) / 1e9
    roof = dict(bound="hbm", kernel="LU-SGS solve (k_scale + 5 x [k_ux, k_rhs, sweeps, k_mid, sweeps, k_fin])",
                achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=None, peak_source=peak_src,
                algorithmic_bytes_per_cell=algo, bytes_per_row_iteration=per_iter, bytes_per_row_setup=setup)
    # e2e: the reference's calling convention -- host arrays in (setELE / setD / setRHSb), host x out
    e2e = None
    if n * (nnz / n) * B * B * 8 < 24e9:
        hv, hb, hx = val.cpu().numpy(), b.cpu().numpy(), x0.cpu().numpy()
        solver.solve(hv, hb, hx, LUSGS_ITERS)
        e0 = time.perf_counter()
        ne = 2
        for _ in range(ne):
            solver.solve(hv, hb, hx, LUSGS_ITERS)
        e2e_s = (time.perf_counter() - e0) / ne
        e2e = dict(value=n / e2e_s, unit="cell-updates/s", h2d_bytes_per_step=int(hv.nbytes + hb.nbytes + hx.nbytes),
                   d2h_bytes_per_step=int(hx.nbytes), ms_per_step=e2e_s * 1e3, steps=ne)
    cpu = None
    if not args.no_cpu:
        cpu = lusgs_cpu(args)
    out = dict(metric="cell_updates_per_sec", value=value, unit="cell-updates/s", n_gpus=1, steps=args.steps,
               warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
               dtype="f64", data="synthetic",
               config=dict(workload=f"lusgs box{args.n}: block-5 LU-SGS (reference SparseSolver::solveILU algebra), "
                                    f"{LUSGS_ITERS} iterations per solve, pattern = tet adjacency of the {args.n}^3 x 6 box "
                                    f"in Hilbert order, {ncol}-colour sweep order, random diagonally dominant blocks",
                           rows=n, blocks=int(nnz), colours=ncol, levels=list(lev),
                           l2="inputs larger than L2" if nnz * 200 > 2 ** 28 else "inputs smaller than L2: flush not applied"),
               clocks=clk, e2e=e2e, gpu_launches=launches, roofline=roof, cpu_baseline=cpu,
               wall_ms_per_step=wall * 1e3 / args.steps, device_gib=dev_bytes / 2 ** 30)
    emit(json.dumps(out))


def lusgs_cpu(args, budget_s=12.0):
    """The oracle's restatement of SparseSolver::solveILU (pinned bit-exactly to the reference build),
    one thread (the reference's sweeps are sequential), on a bounded sample."""
    from oracle import oracle
    _, rowptr, col, order, ncol, dpos = lusgs_system(args, min(args.cpu_n, 48))
    n, nnz, B = rowptr.size - 1, col.size, LUSGS_B
    rng = np.random.default_rng(1)
    val = (rng.random((nnz, B, B)) - 0.5) * 0.2
    val[dpos] += 3.0 * np.eye(B)
    b = rng.random((n, B)); x0 = np.ones((n, B))
    t = time.perf_counter()
    oracle.lusgs(rowptr, col, val, b, x0, B, LUSGS_ITERS)
    t1 = time.perf_counter() - t
    reps = int(min(50, max(1, round(budget_s / max(t1, 1e-3)))))
    t = time.perf_counter()
    for _ in range(reps):
        oracle.lusgs(rowptr, col, val, b, x0, B, LUSGS_ITERS)
    el = time.perf_counter() - t
    return dict(value=n * reps / el, unit="cell-updates/s", cores=1, kind="port",
                sample=f"lusgs box n={min(args.cpu_n, 48)}: {n} rows x {reps} solves of {LUSGS_ITERS} iterations in {el:.2f}s "
                       "(oracle/lusgs_oracle.cpp, sequential like the reference)")


def measure(args, rank, world, dist, local, want_cpu=True):
    """One workload through the CUDA path: K timed steps with the state resident in HBM, digest, every-step-residual
    leg, roofline of the dominant kernel, e2e leg with host buffers.  Returns the JSON line as a dict."""
    import torch
    import mstgpu
    f, Q0, desc = build_workload(args)
    nc_total, U, D = f["ncells"], f["dim"] + 2, f["dim"]
    dt_run, inlet = run_params(args, Q0)  # the inlet of the step / sphere cases is their free stream
    t = time.time()
    if world > 1:
        # one partition per GPU (equal ranges of the Hilbert curve), 2 ghost layers
        part = mstgpu.Partition(f, world, rank, order=2)
        log(f"[bench] rank {rank}: {part.n_owned} owned + {part.n_local - part.n_owned} ghost cells, "
            f"{part.n_neighbors} neighbours, partition in {time.time() - t:.1f}s")
        ctx = mstgpu.Context(part, order=args.order, flux=args.flux, device=local, kernel=args.kernel, inletQ=inlet,
                             viscous=args.viscous, tile_cells=args.tile_cells, block_threads=args.block_threads, renumber=args.renumber,
                             **scheme_kwargs(args))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(mstgpu.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        ctx.comm_init(world, rank, idt.cpu().numpy().tobytes())
        # halo through peer memory (stores into the neighbours' ghost blocks over NVLink + epoch flags) when every
        # rank can open its neighbours' buffers; ncclSend/ncclRecv otherwise (the reductions stay on NCCL)
        halo_mode = "nccl send/recv (pack kernel + grouped ncclSend/ncclRecv)"
        if args.halo == "peer" and ctx.peer_connect_torch(dist, world, rank, device="cuda"):
            halo_mode = "peer-memory push (one kernel stores boundary rows into the neighbours' ghost blocks over NVLink, epoch flags; CUDA IPC)"
        Q0 = np.ascontiguousarray(Q0[part.cell_ids[:part.n_owned]])
        nc = part.n_owned
        del f["cf_idx"]
    else:
        ctx = mstgpu.Context(f, order=args.order, flux=args.flux, device=local, kernel=args.kernel, inletQ=inlet,
                             viscous=args.viscous, tile_cells=args.tile_cells, block_threads=args.block_threads, renumber=args.renumber,
                             **scheme_kwargs(args))
        nc = nc_total
        halo_mode = None
    log(f"[bench] context built in {time.time() - t:.1f}s, {ctx.device_bytes / 2**30:.2f} GiB on device")
    ctx.set_state(Q0)
    if args.implicit:
        # BASELINE config 5: every step = explicit residual + block assembly + 5 LU-SGS sweeps (colour-ordered)
        # with several GPUs: every rank sweeps its own rows, ghost couplings lag one sweep
        t = time.time()
        ctx.implicit_setup(True)
        log(f"[bench] implicit setup in {time.time() - t:.1f}s, {ctx.device_bytes / 2**30:.2f} GiB on device")
        dt_run = args.implicit_dt
        ctx.step = lambda dt, n: ctx.step_implicit(dt, n, LUSGS_ITERS)
    if args.cfl > 0:
        # extension: every step at its own global CFL step, computed on the device (+ allreduce(min))
        ctx.step = lambda dt, n: ctx.step_cfl(args.cfl, n)
    ctx.step(dt_run, args.warmup)
    ctx.sync()
    # ---- timed region: K steps, state resident in HBM -------------------------
    # Explicit fixed-dt steps: ONE mstgpu_step(dt, K) call as a user issues it (pairs of steps from the CUDA graph;
    # with several GPUs the halo push, the flag wait and both tile classes are nodes of that graph), bracketed by
    # CUDA events on the solver's stream.  The per-kernel durations of the roofline come from a second,
    # instrumented pass of the same K steps right after it (events around every launch force the launches out of
    # the graph and serialise the two streams of a partitioned step, so they are kept out of `value`).
    # Implicit and CFL steps have no graph path: one pass, instrumented.
    plain = not args.implicit and args.cfl <= 0
    ctx.enable_kernel_timing(not plain)
    l0 = ctx.launch_count
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.3)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    if args.implicit:
        ms = ctx.step_implicit(dt_run, args.steps, LUSGS_ITERS)  # CUDA events on the solver's own stream
    elif args.cfl > 0:
        ms = ctx.step_cfl_timed(args.cfl, args.steps)  # CUDA events on the solver's own stream
    else:
        ms = ctx.step_timed(dt_run, args.steps)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    wall = time.perf_counter() - w0
    clk = clocks.stop()
    ranks_info = None
    if dist is not None:  # device time of the slowest rank
        mine = torch.tensor([ms, clk.get("sm_mhz") or 0.0, clk.get("power_w_max") or 0.0, clk.get("mem_mhz") or 0.0, clk.get("temp_c_max") or 0.0,
                             float(len(clk.get("reasons") or []))], dtype=torch.float64, device="cuda")
        allm = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allm, mine)
        ranks_info = [dict(rank=i, ms_per_step=round(float(m[0]) / args.steps, 4), sm_mhz=float(m[1]), power_w_max=float(m[2]), mem_mhz=float(m[3]),
                           temp_c_max=float(m[4]), throttle_reasons=int(m[5])) for i, m in enumerate(allm)]
        tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    launches = ctx.launch_count - l0
    res = ctx.residual()
    value = nc_total * args.steps / (ms * 1e-3)
    log(f"[bench] {args.steps} steps in {ms:.2f} ms (wall {wall * 1e3:.2f} ms), residual {res}")

    # ---- state digest after W + K steps: order-independent, keyed by global cell id, summed over ranks.
    # Every N runs the same W + K steps from the same initial state, so equal digests across the N = 1, 2, 4, 8
    # lines are bit-identity of the partitioned runs with the single-GPU run.
    gids = part.cell_ids[:part.n_owned] if world > 1 else np.arange(nc_total)
    dg = state_digest(ctx.get_state(), gids)
    if dist is not None:
        tdg = torch.tensor([dg - (1 << 64) if dg >= (1 << 63) else dg], dtype=torch.int64, device="cuda")
        dist.all_reduce(tdg, op=dist.ReduceOp.SUM)  # two's-complement wrap = sum mod 2^64
        dg = int(tdg.item()) & ((1 << 64) - 1)
    digest = dict(value=f"{dg:016x}", after_steps=args.warmup + args.steps,
                  what="sum mod 2^64 over all cells and variables of splitmix64(bits(Q) ^ golden*(global_cell_id*U+k+1)), all-reduced over ranks")
    log(f"[bench] state digest after {args.warmup + args.steps} steps: {dg:016x}")

    # ---- instrumented pass: the same K steps with CUDA events around every launch ----------------------------
    inst_ms = None
    if plain:
        ctx.enable_kernel_timing(True)
        inst_ms = ctx.step_timed(dt_run, args.steps)
        if dist is not None:
            ti = torch.tensor([inst_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(ti, op=dist.ReduceOp.MAX)
            inst_ms = float(ti.item())
    kt = {k: ctx.kernel_time(k) for k in ("gradient", "gradient_lsq", "limiter", "flux", "update", "step_tiles", "halo_pack",
                                           "assemble", "lusgs", "increment", "halo_exchange", "halo_tiles", "interior_tiles")}
    kt = {k: v for k, v in kt.items() if v[1] > 0}
    ctx.enable_kernel_timing(False)

    # ---- the same K steps with the residual of EVERY step (Time.cpp:69-76 computes it each step; the timed
    # region above is one mstgpu_step(dt, K) call, which reduces the residual of its last step only) ----------
    every = None
    if plain:
        ms1 = 0.0
        for _ in range(args.steps):
            ms1 += ctx.step_timed(dt_run, 1)
        if dist is not None:
            t1 = torch.tensor([ms1], dtype=torch.float64, device="cuda")
            dist.all_reduce(t1, op=dist.ReduceOp.MAX)
            ms1 = float(t1.item())
        every = dict(value=nc_total * args.steps / (ms1 * 1e-3), unit="cell-updates/s", ms_per_step=ms1 / args.steps,
                     what=f"{args.steps} calls of mstgpu_step(dt, 1): residual reduced on every step, one host call per step")

    # ---- roofline of the dominant kernel --------------------------------------
    peak, peak_src = measured_peaks()
    ab = dict(ALGO_BYTES[(D, args.order)])
    if args.order == 2 and args.limiter != "none":
        # limiter pass: Q in, gradient in + out, r = fc - cc per (cell, face), eps^2
        fpc = 2.0 if D == 3 else 1.5
        ab["step"] += 8 * U + 2 * 8 * U * D + 2 * fpc * 8 * D + 8
    per_kernel = {k: (v[0] / max(v[1], 1)) for k, v in kt.items()}
    # spans of the partitioned step on its two streams (rank 0's own): exchange + wait, halo tiles, interior tiles
    breakdown = {k: round(per_kernel.pop(k), 4) for k in ("halo_exchange", "halo_tiles", "interior_tiles") if k in per_kernel}
    breakdown_ranks = None
    if dist is not None and breakdown:
        # every rank's spans (the step time is the slowest rank's; the exchange span of a fast rank is mostly waiting)
        tb = torch.tensor([breakdown.get(k, 0.0) for k in ("halo_exchange", "halo_tiles", "interior_tiles")] + [float(part.n_local - part.n_owned), float(part.n_neighbors)],
                          dtype=torch.float64, device="cuda")
        allb = [torch.zeros_like(tb) for _ in range(world)]
        dist.all_gather(allb, tb)
        breakdown_ranks = [dict(rank=i, halo_exchange=round(float(b[0]), 4), halo_tiles=round(float(b[1]), 4), interior_tiles=round(float(b[2]), 4),
                                ghost_cells=int(b[3]), neighbours=int(b[4])) for i, b in enumerate(allb)]
    dom = max(per_kernel, key=per_kernel.get)
    if args.viscous:
        ab = dict(ab, step=616, flux_update=368) if (D, args.order) == (3, 2) else ab  # + eta per face
    if args.implicit:
        # the LU-SGS solve dominates an implicit step: its own algorithmic bytes over its duration
        nnz_row = 1.0 + 2.0 * (f["nint"] / nc_total)
        per_iter, setup = lusgs_bytes_per_row(nnz_row)
        abytes = LUSGS_ITERS * per_iter + setup
        achieved = abytes * nc / (per_kernel["lusgs"] * 1e-3) / 1e9
        kname = "LU-SGS solve of the implicit step (k_diag, k_scale, 5 x [k_ux, k_rhs, sweeps, k_mid, sweeps, k_fin])"
        ab["step"] = ab["step"] + abytes + nnz_row * 8 * U * U + 2 * 8 * U  # + assembly writes, increment
    elif "step_tiles" in per_kernel:
        # fused kernel: one launch does both passes of SURVEY.md 8(d) -> the whole
        # step's algorithmic bytes (600 B per tet cell-update) over its duration
        achieved = ab["step"] * nc / (per_kernel["step_tiles"] * 1e-3) / 1e9
        kname, abytes = "k_step_tiles (gradient+flux+update fused)", ab["step"]
    else:
        # split path: pass 2 of 8(d) is flux + update; its bytes are charged to the two together
        pass2_ms = per_kernel["flux"] + per_kernel["update"]
        achieved = ab["flux_update"] * nc / (pass2_ms * 1e-3) / 1e9
        kname, abytes = "flux+update (pass 2 of 8d)", ab["flux_update"]
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel: only where a committed ncu
    # capture of THIS instantiation exists (profiles/traffic.json, keyed by dim/order/options); null otherwise
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and "step_tiles" in per_kernel and not args.implicit:
        tj = json.load(open(tp))
        key = f"step_tiles_d{D}_o{args.order}" + ("_visc" if args.viscous else "") + (f"_{args.limiter}" if args.limiter != "none" else "")
        if key in tj:
            traffic, traffic_src = tj[key]["bytes_per_cell"] * nc, f"profiles/traffic.json[{key}]: " + tj[key]["source"]
    roof = dict(bound="hbm", kernel=kname, achieved=achieved, peak=peak, unit="GB/s",
                frac=achieved / peak, traffic=traffic, traffic_source=traffic_src, peak_source=peak_src,
                algorithmic_bytes_per_cell=abytes,
                kernels_ms={k: round(v, 4) for k, v in per_kernel.items()}, dominant=dom,
                step_frac=ab["step"] * value / 1e9 / peak)

    # ---- e2e: reference-facing call with HOST buffers -----------------------------
    # Time.cpp:63-76 runs solve() and then reads old and new state on the host: one mstgpu_step_host call per
    # step (H2D of the input rows, the step, D2H of the output rows, pipelined over chunks of host rows) and the
    # residual.  The same sequence as three separate calls (set_state, step, get_state) is timed beside it.
    hin = torch.empty((nc, U), dtype=torch.float64, pin_memory=True)
    hout = torch.empty((nc, U), dtype=torch.float64, pin_memory=True)
    ne = max(2, min(args.steps, 5))

    def e2e_leg(streamed):
        nonlocal hin, hout
        hin.numpy()[:] = Q0
        for it in range(1 + ne):
            if it == 1:
                torch.cuda.synchronize()
                if dist is not None:
                    dist.barrier()
                e0 = time.perf_counter()
            if streamed:
                ctx.step_host(hin.data_ptr(), hout.data_ptr(), dt_run, args.host_chunks)
            else:
                ctx.set_state_ptr(hin.data_ptr())      # H2D of the step's input state
                ctx.step(dt_run, 1)
                ctx.get_state_ptr(hout.data_ptr())     # D2H of the new state (Time.cpp:66-67)
            ctx.residual()                         # D2H of the residual (Time.cpp:69-76)
            hin, hout = hout, hin                  # updateNewToOld on the host side
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        s_ = (time.perf_counter() - e0) / ne
        if dist is not None:
            te = torch.tensor([s_], dtype=torch.float64, device="cuda")
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            s_ = float(te.item())
        return s_, int(hin.numpy().view(np.uint64).sum(dtype=np.uint64))  # checksum of the final state's bit patterns

    streamed_ok = args.kernel == "tiles" and not args.cfl and not args.implicit
    e2e_3, sha3 = e2e_leg(False)
    e2e_s, sha_s = e2e_leg(True) if streamed_ok else (e2e_3, sha3)
    assert sha3 == sha_s, "streamed step differs from set_state + step + get_state"
    e2e = dict(value=nc_total / e2e_s, unit="cell-updates/s", h2d_bytes_per_step=nc * U * 8,
               d2h_bytes_per_step=nc * U * 8 + U * 8, ms_per_step=e2e_s * 1e3, steps=ne,
               call="mstgpu_step_host (H2D, step, D2H pipelined over %d chunks of host rows) + mstgpu_residual_linf" % (args.host_chunks or 64)
                    if streamed_ok else "mstgpu_set_state + mstgpu_step + mstgpu_get_state + mstgpu_residual_linf",
               three_calls=dict(value=nc_total / e2e_3, ms_per_step=e2e_3 * 1e3,
                                what="mstgpu_set_state + mstgpu_step(dt, 1) + mstgpu_get_state + mstgpu_residual_linf, one after the other"))

    cpu = cpu_baseline(args) if (world == 1 and not args.no_cpu and want_cpu) else None

    out = dict(metric="cell_updates_per_sec", value=value, unit="cell-updates/s", n_gpus=world,
               steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True,
               scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
               config=workload_config(args, world, nc_total, f["nfaces"], U, desc, dt_run),
               gpu_config=dict(kernel=args.kernel, tile_cells=args.tile_cells, renumber=args.renumber, block_threads=args.block_threads,
                               halo=halo_mode, per_rank=ranks_info, step_breakdown_ms=breakdown or None, step_breakdown_ms_per_rank=breakdown_ranks,
                               timed_call="one mstgpu_step(dt, K) call: pairs of steps from the CUDA graph, the last step(s) launched directly"
                                          if plain else "K steps, CUDA events around every launch",
                               instrumented_ms_per_step=None if inst_ms is None else inst_ms / args.steps,
                               residual="reduced on the last step of the timed mstgpu_step(dt, K) call (SURVEY 8f.1: residual every k); "
                                        "see every_step_residual for one call per step"),
               clocks=clk, e2e=e2e, gpu_launches=launches, roofline=roof, cpu_baseline=cpu,
               state_digest=digest, every_step_residual=every,
               wall_ms_per_step=wall * 1e3 / args.steps, device_gib=ctx.device_bytes / 2 ** 30)
    if dist is not None:
        dist.barrier()
    ctx.close()
    return out


# BASELINE.json configs 1, 2, 3, 5 next to the headline (config 4): short runs of the same code path, each with its
# own roofline; the bound is named for what it is (configs 1 and 2 fit the 126 MB L2: launch / latency, not HBM)
SECONDARY = [
    ("config1_sod", dict(workload="sod", n=203, order=2, flux="roe", graph=1), 200,
     "launch/latency: 18 282 cells, state + tables live in L2; the HBM fraction is not the binding bound"),
    ("config2_step445_ausm", dict(workload="step", n=445, order=1, flux="ausm", graph=1), 200,
     "L2/latency: 998 046 triangles, state + packets (~0.3 GB) mostly L2-resident, ~60 us per step"),
    ("config3_sphere_roe_viscous", dict(workload="sphere", n=42, order=2, flux="roe", viscous=1), 20, "hbm"),
    ("config5_implicit_lusgs", dict(workload="box", n=203, order=2, flux="roe", implicit=1), 3, "hbm"),
]


def run_ours(args, rank, world):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = measure(args, rank, world, dist, local)
    headline = (args.workload == "box" and args.n == 203 and not args.implicit and not args.viscous and args.limiter == "none"
                and args.cfl <= 0 and args.order == 2 and args.flux == "roe")
    if args.secondary == 1 or (args.secondary < 0 and headline):
        sec = {}
        for name, over, steps, bound in SECONDARY:
            a = argparse.Namespace(**vars(args))
            for k, v in over.items():
                setattr(a, k, v)
            a.steps, a.warmup, a.no_cpu = steps, 5, True  # >= 4 warm-up steps also run the CUDA graph of the launch-bound cases once
            if (name in ("config1_sod", "config2_step445_ausm") and world > 1) or (name == "config3_sphere_roe_viscous" and world > 2):
                continue   # BASELINE.json: configs 1 and 2 on one B200, config 3 on 1 and 2
            t0 = time.time()
            try:
                o = measure(a, rank, world, dist, local, want_cpu=False)
                o["roofline"]["bound_named"] = bound
                sec[name] = {k: o[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "config", "gpu_config", "roofline",
                                               "e2e", "gpu_launches", "state_digest", "device_gib", "n_gpus")}
            except BaseException as e:  # a secondary line must never cost the headline
                sec[name] = dict(error=f"{type(e).__name__}: {e}"[:300])
                if dist is not None:
                    raise
            log(f"[bench] secondary {name}: {time.time() - t0:.1f}s")
        out["secondary"] = sec
    if rank == 0:
        emit(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: str):
    """the ONE JSON line, on the process's real stdout"""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (line + "\n").encode())


def main():
    # stdout carries exactly one JSON line.  Native libraries write banners to fd 1 behind Python's back (NCCL's
    # version line at any NCCL_DEBUG level >= VERSION): fd 1 is pointed at stderr for the whole run and the JSON
    # line goes to a saved duplicate of the real stdout.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="box", choices=["box", "step", "sphere", "lusgs", "msh", "sod"])
    ap.add_argument("--mesh", default="", help="--workload msh: a Fluent .msh file (the subset the reference's reader accepts)")
    ap.add_argument("--size", "--n", dest="n", type=int, default=203, help="hexes per side (box) or 1/h (step)")
    ap.add_argument("--cpu-n", type=int, default=96, help="size of the bounded CPU sample of the GPU arm's cpu_baseline leg")
    ap.add_argument("--cpu-sample", type=int, default=0, help="--impl reference: run an N^3 sample instead of the full workload (0 = full size)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--kernel", default="tiles", choices=["tiles", "split"])
    ap.add_argument("--tile-cells", type=int, default=0)
    ap.add_argument("--renumber", type=int, default=2, help="0 none, 1 Morton, 2 Hilbert")
    ap.add_argument("--flux", default="roe", choices=["roe", "ausm"])
    ap.add_argument("--order", type=int, default=2, choices=[1, 2])
    ap.add_argument("--viscous", type=int, default=0, choices=[0, 1], help="laminar viscous term (extension; fused kernel at second order)")
    ap.add_argument("--block-threads", type=int, default=0)
    ap.add_argument("--gradient", default="gg", choices=["gg", "lsq"], help="extension: least-squares gradient")
    ap.add_argument("--limiter", default="none", choices=["none", "bj", "venkat"], help="extension: slope limiter")
    ap.add_argument("--limiter-k", type=float, default=5.0)
    ap.add_argument("--cfl", type=float, default=0.0, help="extension: > 0 = global CFL time step instead of DT")
    ap.add_argument("--shock", type=int, default=0, choices=[0, 1],
                    help="box init: 1 = SOD split at x = 0.5 + perturbation (SURVEY 8d; needs --limiter and --cfl)")
    ap.add_argument("--implicit", type=int, default=0, choices=[0, 1],
                    help="config 5: implicit steps (block assembly + 5 colour-ordered LU-SGS sweeps)")
    ap.add_argument("--implicit-dt", type=float, default=1e-3)
    ap.add_argument("--graph", type=int, default=0, choices=[0, 1], help="also time the K steps issued from the CUDA graph")
    ap.add_argument("--host-chunks", type=int, default=0, help="chunks of host rows of the streamed e2e step (0 = library default, 64)")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"], help="N > 1: halo exchange through peer memory (default) or NCCL send/recv")
    ap.add_argument("--secondary", type=int, default=-1, help="BASELINE configs 1, 2, 3, 5 as short runs under the `secondary` key: "
                    "1 = always, 0 = never, -1 = with the headline workload only (default)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
    elif args.workload == "lusgs":
        run_lusgs(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
