#!/usr/bin/env python
"""Benchmark of the STARDIS opacity + formal-solution hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own numba CPU path on the host cores
    python bench.py --workload {sim10aa,sim100aa,solar_full,solar_weak,astar,coolgiant_ir,grid_sweep64}

metric  : emergent-spectrum nu-points/sec (opacity + raytrace); a "step" is ONE pass of the hot path over the whole
          frequency grid of the workload: broadening (K1) -> line windows/records -> Voigt accumulation (K2) ->
          continuum + total (K3) -> formal solution for all angles (K4).
workload: default = BASELINE.json configs[1]: solar MARCS structure, 3000-10000 A at 0.01 A (N = 700 000), D = 56, 10
          angles, all four broadening mechanisms, H- bf table + H I bf/ff + electron scattering, SYNTHETIC line list of
          300 000 lines (real Kurucz/CD23 data does not exist offline; recipe in SURVEY.md 8d / plasma/synthetic.py).
          The other names are the remaining BASELINE configs (and a weak-line regime of the flagship).
value   : N / (device time of one step), inputs resident in HBM, CUDA events on the launching stream, max over ranks.
e2e     : same metric through the public API (calc_alphas + raytrace + read of the emergent spectrum) with HOST
          (pinned) inputs: H2D of the line table / grid and D2H of the spectrum inside the timed region.
parity  : the GPU's total opacity and flux on a sampled nu shard against the CPU oracle evaluated on the same shard in
          the same run (rtol 1e-8 / 1e-6, BASELINE north_star); the process exits non-zero when it fails.
roofline: per kernel from EXECUTED work (counting instantiation of the kernels) and the kernel's device time measured
          live with CUDA events inside the library (sd_phase_times), against the FP64 peak measured in the same run.
N > 1   : the frequency grid is sharded across ranks (strong scaling; global windows, no exchange inside the
          kernels); the emergent spectrum is all-gathered with NCCL inside the timed region.  grid_sweep64 instead
          scatters 64 independent stellar models over the ranks (replicas, spectra gathered).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "emergent-spectrum nu-points/sec (opacity+raytrace)"
FLOPS_PER_EVAL = np.array([19.0, 33.0, 65.0, 164.0])  # Humlicek regions I..IV, FMA = 2 flops
BYTES_PER_CELL = 16.0                                  # write total_alphas + write F_nu
# executed work of the far-field path, FP64 flops (FMA = 2), counted from the source (csrc/k2_lines.cu):
FLOPS_FAR_SETUP = 38.0   # per (pair, tile) expansion: pole distances, two reciprocals, recurrence constants
FLOPS_FAR_TERM = 9.0     # per Taylor term: coefficient FMA + add, two three-term recurrences (FMA + MUL each)
FLOPS_HORNER = 65.0      # per (pixel, depth, level): 31 Horner FMAs (SD_FAR_K) + scaled argument + accumulate
FAR_LEVELS = 4           # SD_FAR_LEVELS
FLOPS_S2M = 38.0 + 32 * 9.0   # per (pair, level) multipole expansion: set-up + 32 moments (same recurrence as a Taylor term)
FLOPS_M2L_STEP = 32 * 2.0 + 2.0 * 32 / 14   # per (source, target, depth, k) row step: one FMA on 32 lanes (+ the row update shared by 14 depths)
ALPHA_RTOL, F_RTOL = 1e-8, 1e-6


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="solar_full")
    ap.add_argument("--lines", type=int, default=None, help="override the number of synthetic lines")
    ap.add_argument("--partition", default="depth", choices=["depth", "nu"],
                    help="multi-GPU decomposition of the opacity stages (the formal solution is always sharded by nu)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-direct", action="store_true", help="skip the direct-mode (far field off) comparison run")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of one reference sample")
    ap.add_argument("--cpu-kind", default="reference", choices=["reference", "port"],
                    help="CPU leg: the reference's numba path (oracle/_ref) or the C port (oracle/)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- workload
def opacity_config(table_dir, n_thetas):
    from stardis_b200.io.config import Configuration, validate_config
    from stardis_b200.synthetic import write_cross_section_files

    paths = write_cross_section_files(table_dir)
    raw = dict(stardis_config_version=1.0, atom_data="synthetic:0", input_model=dict(type="marcs", fname="sun.mod"),
               opacity=dict(file={"Hminus_bf": paths["Hminus_bf"]}, bf={"H_I": {}}, ff={"H_I": {}},
                            disable_electron_scattering=False,
                            line=dict(disable=False, broadening=["radiation", "linear_stark", "quadratic_stark", "van_der_waals"])),
               no_of_thetas=n_thetas)
    return Configuration(validate_config(raw))


def build_workload(args, name=None):
    from stardis_b200.synthetic import WORKLOADS, make_workload

    name = name or args.workload
    w = make_workload(name, seed=1, n_lines=args.lines, device_strengths=True)
    cfg = opacity_config(tempfile.mkdtemp(prefix="sdb200_tables_"), WORKLOADS[name][5])
    desc = (f"{name}: MARCS {WORKLOADS[name][0]} structure (T x {WORKLOADS[name][1]:.3f}) D={len(w['atmosphere']['T'])}, lambda "
            f"{w['lambdas'].value[0]:.0f}-{w['lambdas'].value[-1]:.0f} A step 0.01 (N={len(w['nus'])}), "
            f"{len(w['plasma']._line_table)} synthetic lines, {cfg.no_of_thetas} angles, 4 broadenings, "
            f"Hminus_bf+H_I bf/ff+e- continuum")
    return w, cfg, desc


def config_block(desc, n_lines, D, cells, world, bounds, partition="depth"):
    if world > 1 and partition == "depth":
        part = (f"opacity stages: depth points r, r+{world}, ... of the whole grid on rank r; one NCCL all-to-all of the total "
                f"opacity; formal solution: {world} equal-width nu ranges; all-gather of the spectrum")
    else:
        part = f"{world} contiguous nu range(s), global windows, cut at equal (pixels + 10 x lines inside)"
    return {"workload": desc,
            "l2": "inputs larger than L2 (line records %.2f GB, outputs %.2f GB per array)" % (n_lines * D * 64 / 1e9, cells * 8 / 1e9),
            "partition": part, "ranges": [list(b) for b in bounds]}


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            tok = [t.strip() for t in ln.split(",")]
            if len(tok) < 6:
                continue
            try:
                sm.append(float(tok[0]))
                smax = float(tok[1])
            except ValueError:
                continue
            for name, val in zip(names, tok[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU reference arm
def cpu_block(r):
    out = {"value": r["value"], "unit": "nu-points/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    if "full_step_seconds" in r:
        out["full_step_seconds"] = r["full_step_seconds"]
    return out


def run_reference_arm(args):
    """The reference's own CPU implementation of the path (its numba-parallel functions, executed unmodified from
    oracle/_ref through oracle/ref_shim.py; the C port when numba or the staged modules are missing -- `kind` says
    which), all host cores, each step one bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_leg

    name = "solar_full" if args.workload == "grid_sweep64" else args.workload
    w, cfg, desc = build_workload(args, name)
    vals, r = [], None
    for i in range(args.warmup + args.steps):
        secs = args.cpu_seconds if i >= args.warmup else min(args.cpu_seconds, 3.0)
        r = ref_leg.cpu_sample(w, cfg, secs, prefer=args.cpu_kind)
        if i >= args.warmup:
            vals.append((r["value"], r["seconds"]))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([t for _, t in vals]) * 1e3)
    N, D, L = len(w["nus"]), w["model"].no_of_depth_points, len(w["plasma"]._line_table)
    blk = cpu_block(r)
    blk["value"] = value
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "nu-points/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_block(desc, L, D, D * N, 1, [(0, N)]), "cpu_baseline": blk,
           "e2e": {"value": value, "unit": "nu-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------- B200 arm
def rel_err(a, b):
    """max |a - b| / |b| with 0/0 = 0 (np.testing.assert_allclose semantics, atol = 0)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    diff = np.abs(a - b)
    den = np.abs(b)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(diff == 0.0, 0.0, diff / den)
    return float(np.nanmax(r)) if r.size else 0.0


class HotPath:
    """Resident inputs of one workload on one rank + the device step (pointers into the C ABI).

    world == 1: one context, all depth points, whole grid.  world > 1 (``partition`` "depth"): the opacity stages
    (K1, preparation, K2, K3) run on ``ctx_op`` for the depth points rank, rank + world, ... of the WHOLE grid, one
    all-to-all moves the total opacity to the equal-width pixel ranges, the formal solution runs on ``ctx`` for all
    depth points of this rank's range, the emergent spectrum is all-gathered.  ``partition`` "nu": everything on this
    rank's cost-balanced pixel range (no exchange; the per-(line, depth) preparation is repeated on every rank)."""

    def __init__(self, ctx, w, cfg, rank, world, torch, partition="depth", stream=None):
        from stardis_b200 import _lib as L
        from stardis_b200 import units as u
        from stardis_b200.device import DeviceContext
        from stardis_b200.distributed import all_shards, depth_indices, line_balanced_bounds
        from stardis_b200.radiation_field import RadiationField
        from stardis_b200.radiation_field.opacities.opacities_solvers import base as ob
        from stardis_b200.radiation_field.radiation_field_solvers.base import ray_distances

        self.L_, self.ctx, self.torch, self.world, self.rank = L, ctx, torch, world, rank
        model, plasma, nus = w["model"], w["plasma"], w["nus"]
        self.model, self.plasma, self.nus, self.cfg = model, plasma, nus, cfg
        self.N, self.D = len(nus), model.no_of_depth_points
        self.depth_mode = world > 1 and partition == "depth"
        line_cfg = cfg.opacity.line
        self.nus_q = u.Quantity(nus, u.Hz)
        self.lines = ob.select_lines(plasma, model, self.nus_q, line_cfg)
        if self.depth_mode:
            self.bounds = all_shards(self.N, world)
            self.didx = depth_indices(self.D, rank, world)
            self.ctx_op = DeviceContext(ctx.device, stream=stream)
        else:
            # cost-balanced contiguous nu ranges (pixels + 10 x lines inside): the blue end of the grid holds ~10x more
            # lines per pixel than the red end, so equal-width ranges would leave rank 0 with several times the core work
            self.bounds = line_balanced_bounds(nus, self.lines.nu, world)
            self.didx = np.arange(self.D)
            self.ctx_op = ctx
        self.p0, self.p1 = self.bounds[rank]
        self.W = self.p1 - self.p0
        self.flags = ob._line_flags(line_cfg)
        didx = self.didx
        tables, _ = ob.file_tables(plasma, model, cfg.opacity.file)
        self.tables = [dict(t, depth_scale=np.ascontiguousarray(t["depth_scale"][didx]),
                            depth_y=None if t.get("depth_y") is None else np.ascontiguousarray(t["depth_y"][didx])) for t in tables]
        self.bf_cut, bf_prefix = ob.bf_descriptor(plasma, cfg.opacity.bf)
        self.bf_prefix = np.ascontiguousarray(bf_prefix[:, didx])
        self.ff_coef = np.ascontiguousarray(ob.ff_descriptor(plasma, model, cfg.opacity.ff)[didx])
        self.rayleigh = tuple(np.ascontiguousarray(c[didx]) for c in ob.rayleigh_descriptor(plasma, model, cfg.opacity.rayleigh))
        self.electron = np.ascontiguousarray(ob.electron_descriptor(plasma)[didx])
        self.srf0 = RadiationField(self.nus_q, None, model, cfg.no_of_thetas)
        self.ds, self.inward = ray_distances(model, self.srf0.thetas)
        T = u.values_of(model.temperatures)
        n_e = np.asarray(plasma.electron_densities.values, dtype=np.float64)
        n_H = np.asarray(plasma.ion_number_density.loc[1, 0].values, dtype=np.float64)
        self.model_vmic = float(u.cgs_values_of(model.microturbulence))
        self.ctx_op.set_atmosphere(T[didx], n_e[didx], n_H[didx], self.model_vmic)
        if self.depth_mode:
            self.ctx_op.set_grid(nus)
            ctx.set_atmosphere(T, n_e, n_H, self.model_vmic)
            self.t_local = torch.empty((len(didx), self.N), dtype=torch.float64, device="cuda")
        ctx.set_grid(nus, self.p0, self.p1)
        cols = {k: np.ascontiguousarray(getattr(self.lines, k)) for k in
                ("nu", "mass", "atomic_number", "ion_number", "ionization_energy", "level_energy_upper", "level_energy_lower", "A_ul")}
        cols["alpha_line"] = np.ascontiguousarray(self.lines.alpha_line[:, didx])
        self.d_lines = {k: torch.from_numpy(v).cuda() for k, v in cols.items()}
        self.d_spec = torch.empty(self.W, dtype=torch.float64, device="cuda")

    N_EVENTS = 7
    PHASES = ("K1_broadening", "K2_prepare_and_lines", "K3_continuum", "depth_to_nu_exchange", "K4_raytrace", "spectrum_gather")

    def step(self, ev=None, stream=None, gather=True):
        """One pass of the hot path on resident inputs."""
        from stardis_b200.distributed import allgather_spectrum, exchange_depth_to_nu

        ctx, op, d, L = self.ctx, self.ctx_op, self.d_lines, self.L_
        rec = (lambda i: ev[i].record(stream)) if ev is not None else (lambda i: None)
        rec(0)
        op.set_lines(d["nu"], d["alpha_line"], mass=d["mass"], atomic_number=d["atomic_number"], ion_number=d["ion_number"],
                     ionization_energy=d["ionization_energy"], level_energy_upper=d["level_energy_upper"],
                     level_energy_lower=d["level_energy_lower"], A_ul=d["A_ul"])
        op.calc_broadening(self.flags)                       # K1
        rec(1)
        op.calc_alpha_line(0)                                # K2 (+ window/record preparation)
        rec(2)
        op.calc_continuum(bf_nu_cut=self.bf_cut, bf_prefix=self.bf_prefix, ff_coef=self.ff_coef, rayleigh=self.rayleigh,
                          electron=self.electron, tables=self.tables, store_mask=0)  # K3
        rec(3)
        if self.depth_mode:                                  # (D/R, N) depth rows -> (D, N/R) pixel range, NCCL all-to-all
            op.get(L.BUF_TOTAL, out=self.t_local)
            ctx.set_total(exchange_depth_to_nu(self.t_local, self.D, self.N, bounds=self.bounds))
        rec(4)
        ctx.raytrace(self.ds, self.srf0.I_nus_weights, inward_rays=self.inward)  # K4
        rec(5)
        ctx.get_row(L.BUF_F_NU, -1, out=self.d_spec)
        if self.world > 1 and gather:
            allgather_spectrum(self.d_spec, (self.p0, self.p1), self.N, bounds=self.bounds)
        rec(6)


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from stardis_b200 import _lib as L
    from stardis_b200 import units as u
    from stardis_b200.device import DeviceContext
    from stardis_b200.distributed import allgather_spectrum, stripe_rows
    from stardis_b200.radiation_field import RadiationField
    from stardis_b200.radiation_field.opacities.opacities_solvers import base as ob
    from stardis_b200.radiation_field.radiation_field_solvers.base import raytrace

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()  # all kernels, copies and timing events of this benchmark live on this stream
    torch.cuda.set_stream(stream)
    ctx = DeviceContext(local_rank, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.workload == "grid_sweep64":
        return run_grid_sweep(args, ctx, stream, rank, world, local_rank, barrier, max_over_ranks)

    w, cfg, desc = build_workload(args)
    hp = HotPath(ctx, w, cfg, rank, world, torch, partition=args.partition, stream=stream)
    model, plasma, nus, N, D, W, p0, p1, bounds = hp.model, hp.plasma, hp.nus, hp.N, hp.D, hp.W, hp.p0, hp.p1, hp.bounds
    op = hp.ctx_op   # context of the opacity stages (== ctx unless the multi-GPU run shards them by depth)
    line_cfg = cfg.opacity.line
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(hp.N_EVENTS)]

    sampler = ClockSampler(local_rank)  # sampled from the warm-up on: the timed region itself may last < 1 s
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        hp.step()
    barrier()
    launches0 = ctx.launch_count()
    phase = np.zeros(hp.N_EVENTS - 1)
    kern = {}
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record(stream)
    for _ in range(args.steps):
        hp.step(ev, stream)
        ev[-1].synchronize()
        ctx._keep.clear()
        op._keep.clear()
        phase += [ev[i].elapsed_time(ev[i + 1]) for i in range(hp.N_EVENTS - 1)]
        pt = dict(op.phase_times())  # per-kernel device times of this step (events inside the library)
        pt["K4_raytrace"] = ctx.phase_times()["K4_raytrace"]
        for k, v in pt.items():
            kern[k] = kern.get(k, 0.0) + max(v, 0.0)
    t_end.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    ms_per_step = max_over_ranks(t_start.elapsed_time(t_end)) / args.steps
    phase /= args.steps
    kern = {k: v / args.steps for k, v in kern.items()}
    value = N / (ms_per_step * 1e-3)

    # ---- K2 statistics (untimed): executed work of the default mode + reference-equivalent evaluation counts
    op.set_line_stats(True)
    op.calc_alpha_line(0)
    stats = op.line_stats()
    ex = op.line_stats_ex()
    op.set_line_stats(False)
    reg = torch.tensor(stats["region_evals"].astype(np.float64), device="cuda")
    if world > 1:
        dist.all_reduce(reg)
    region_evals = reg.cpu().numpy()

    def time_k2(reps):
        """Device time of the Voigt accumulation alone (k_far_coeffs + k_lines; records already prepared)."""
        op.calc_alpha_line(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = []
        for _ in range(reps):
            e0.record(stream)
            op.lib.sd_calc_alpha_line(op.h, 0)
            e1.record(stream)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.mean(ms))

    k2_far_ms = time_k2(max(2, args.steps))
    dfma_peak = ctx.bench_dfma(8192)

    # ---- the same step with the far-field expansion switched OFF: every (line, depth, pixel) triple is evaluated
    # directly, like the reference's loop ("roofline_direct").
    direct = None
    if not args.no_direct:
        a_far = torch.empty((len(hp.didx), op.W), dtype=torch.float64, device="cuda")
        op.get(L.BUF_ALPHA_LINE, out=a_far)
        op.set_farfield(False)
        hp.step()
        barrier()
        n_direct = max(1, args.steps // 3)
        td0, td1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        td0.record(stream)
        for _ in range(n_direct):
            hp.step()
        td1.record(stream)
        barrier()
        ms_direct = max_over_ranks(td0.elapsed_time(td1) / n_direct)
        k2_kernel_ms = time_k2(max(1, n_direct))
        a_dir = torch.empty_like(a_far)
        op.get(L.BUF_ALPHA_LINE, out=a_dir)
        torch.cuda.synchronize()
        far_dev = max_over_ranks(float(((a_far - a_dir).abs() / a_dir.abs().clamp_min(1e-300)).max().item()))
        del a_far, a_dir
        op.set_farfield(True)
        flops_direct = float((stats["region_evals"] * FLOPS_PER_EVAL).sum())  # this rank's launch
        direct = dict(ms_per_step=ms_direct, k2_ms=k2_kernel_ms, far_dev=far_dev,
                      tflops=flops_direct / (k2_kernel_ms * 1e-3) / 1e12)
    hp.step()  # default mode again: the buffers hold this rank's total opacity and flux
    barrier()

    # ---- parity on the benched workload: GPU total opacity / flux on a sampled shard vs the CPU oracle (rank 0)
    parity, port = None, None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import ref_leg

        port = ref_leg.port_sample(w, cfg, args.cpu_seconds if args.cpu_kind == "port" else min(args.cpu_seconds, 8.0),
                                   centre=(p0 + p1) // 2, limits=(p0, p1))
        s0, s1 = port["p0"] - p0, port["p1"] - p0
        g_total = torch.empty((D, W), dtype=torch.float64, device="cuda")
        ctx.get(L.BUF_TOTAL, out=g_total)
        a_gpu = g_total[:, s0:s1].cpu().numpy()
        ctx.get(L.BUF_F_NU, out=g_total)
        f_gpu = g_total[:, s0:s1].cpu().numpy()
        del g_total
        parity = {"alpha_max_rel": rel_err(a_gpu, port["total"]), "F_max_rel": rel_err(f_gpu, port["F"]),
                  "pixels": [int(port["p0"]), int(port["p1"])], "rtol_alpha": ALPHA_RTOL, "rtol_F": F_RTOL,
                  "checker": "oracle/stardis_oracle.c on the same shard, same run"}
        parity["ok"] = bool(parity["alpha_max_rel"] <= ALPHA_RTOL and parity["F_max_rel"] <= F_RTOL)

    # ---- end to end through the public API with pinned host inputs
    lt = plasma._line_table
    for name in ("nu", "alpha_line", "atomic_number", "ion_number", "ionization_energy", "level_energy_upper",
                 "level_energy_lower", "A_ul", "mass"):
        a = getattr(lt, name)
        if a is None:
            continue
        pinned = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        setattr(lt, name, pinned.numpy())
        lt.__dict__.setdefault("_pins", []).append(pinned)
    if lt.strength is not None:
        for k, v in list(lt.strength.per_line.items()):
            if v is not None:
                pinned = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
                lt.strength.per_line[k] = pinned.numpy()
                lt.__dict__["_pins"].append(pinned)
    lt._no_autoion = None
    pinned_nus = torch.from_numpy(nus.copy()).pin_memory()
    nus_host = u.Quantity(pinned_nus.numpy(), u.Hz)
    h_spec = torch.empty(W, dtype=torch.float64).pin_memory()
    sel = ob.select_lines(plasma, model, nus_host, line_cfg)
    # per rank: grid + per-line columns + atmosphere + the O(L) inputs of the device line-strength producer
    # (sd_calc_alpha_line_vald fills the (L, D) table in HBM; without producer inputs the table itself travels, striped
    # over the ranks and exchanged over NVLink: distributed.upload_rows_striped)
    r0, r1, _ = stripe_rows(len(sel), rank, world)
    if hp.depth_mode and sel.strength is None:
        r0, r1 = 0, -(-len(sel) * len(hp.didx) // D)  # this rank's depth columns of the (L, D) table
    h2d = int(pinned_nus.numel() * 8 + sum(np.asarray(getattr(sel, k)).nbytes for k in
              ("nu", "mass", "atomic_number", "ion_number", "ionization_energy", "level_energy_upper",
               "level_energy_lower", "A_ul")) + 3 * D * 8 +
              (sel.strength.nbytes() if sel.strength is not None else (r1 - r0) * D * 8))
    d2h = int(W * 8)

    def api_step():
        srf = RadiationField(nus_host, None, model, cfg.no_of_thetas, device_context=ctx, shard=(p0, p1) if world > 1 else None,
                             shard_bounds=bounds, depth_shard=(rank, world) if hp.depth_mode else None)
        ob.calc_alphas(plasma, model, srf, cfg.opacity, store_components=False)
        raytrace(model, srf)
        ctx.get_row(L.BUF_F_NU, -1, out=h_spec)
        ctx.synchronize()
        spec = h_spec
        if world > 1:
            spec = allgather_spectrum(h_spec.numpy(), (p0, p1), N, device="cuda", bounds=bounds)
        return spec

    api_step()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(1, min(args.steps, 5))
    for _ in range(n_e2e):
        api_step()
    barrier()
    e2e_value = N / max_over_ranks((time.perf_counter() - t0) / n_e2e)

    rc = 0
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import ref_leg

            if args.cpu_kind == "reference" and ref_leg.numba_available():
                cpu = cpu_block(ref_leg.numba_sample(w, cfg, args.cpu_seconds))
                cpu["port"] = cpu_block(port)  # the C port, timed on the parity shard in the same run
            else:
                cpu = cpu_block(port)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        if os.path.exists(peaks_path):
            hbm_peak, hbm_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json"
        cells = D * W
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures of this
        # workload (profiles/ncu_traffic.json, written by tools/ncu_summary.py runs); null for any other configuration
        traffic, pipe_busy = {}, {}
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath) and world == 1:
            t = json.load(open(tpath))
            if (t.get("N"), t.get("D"), t.get("L")) == (N, D, len(sel)):
                traffic = t.get("bytes_per_launch", {})
                pipe_busy = t.get("fp64_pipe_busy", {})
        # executed work of this rank's launches in the default (far-field) mode
        flops_lines = float((ex["direct_region_evals"] * FLOPS_PER_EVAL).sum()) + FLOPS_HORNER * FAR_LEVELS * len(hp.didx) * op.W
        flops_far = (FLOPS_FAR_SETUP * ex["far_expansions"] + FLOPS_FAR_TERM * ex["far_terms"]
                     + FLOPS_S2M * ex["multipole_expansions"] + FLOPS_M2L_STEP * ex["m2l_row_steps"])
        t_lines, t_far = kern.get("K2_lines", 0.0), kern.get("K2_far_coeffs", 0.0)
        tf_lines = flops_lines / max(t_lines * 1e-3, 1e-12) / 1e12
        tf_far = flops_far / max(t_far * 1e-3, 1e-12) / 1e12
        t_hbm = kern.get("K3_continuum", 0.0) + kern.get("K4_raytrace", 0.0)
        out = {
            "metric": METRIC, "value": value, "unit": "nu-points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(desc, len(sel), D, cells, world, bounds, args.partition),
            "roofline": {"kernel": "k_lines, default (far-field) mode: the dominant kernel of the timed step", "bound": "fp64",
                         "achieved": tf_lines, "peak": dfma_peak, "unit": "TFLOP/s", "frac": tf_lines / dfma_peak,
                         "traffic": traffic.get("k_lines"), "fp64_pipe_busy_ncu": pipe_busy.get("k_lines"),
                         "peak_source": "measured in this run: dependent-free DFMA loop on all SMs (sd_bench_dfma)",
                         "work": "EXECUTED Voigt evaluations per Humlicek region x SURVEY 8d flops + Horner evaluation of "
                                 "the four far-field polynomials per (pixel, depth)",
                         "flops_per_eval": FLOPS_PER_EVAL.tolist(), "executed_region_evals": ex["direct_region_evals"].tolist(),
                         "flops_horner_per_cell_level": FLOPS_HORNER, "kernel_ms": t_lines,
                         "share_of_step": t_lines / ms_per_step},
            "roofline_far": {"kernel": "k_far_coeffs (+ k_far_reduce) + k_s2m + k_m2l, all four levels", "bound": "fp64", "achieved": tf_far,
                             "peak": dfma_peak, "unit": "TFLOP/s", "frac": tf_far / dfma_peak, "traffic": traffic.get("k_far_coeffs"),
                             "fp64_pipe_busy_ncu": {k: pipe_busy.get(k) for k in ("k_far_coeffs", "k_m2l", "k_m2m", "k_s2m")},
                             "work": "EXECUTED direct (pair, tile) expansions x setup flops + Taylor terms x flops per term + multipole "
                                     "expansions (pair, level) + tile-to-tile translation row steps",
                             "expansions": ex["far_expansions"], "terms": ex["far_terms"],
                             "multipole_expansions": ex["multipole_expansions"], "m2l_row_steps": ex["m2l_row_steps"],
                             "flops_setup": FLOPS_FAR_SETUP, "flops_per_term": FLOPS_FAR_TERM, "kernel_ms": t_far,
                             "share_of_step": t_far / ms_per_step},
            "roofline_hbm": {"kernel": "k_continuum + k_raytrace (K3+K4)", "bound": "hbm",
                             "achieved": BYTES_PER_CELL * cells / max(t_hbm * 1e-3, 1e-12) / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "frac": BYTES_PER_CELL * cells / max(t_hbm * 1e-3, 1e-12) / 1e9 / hbm_peak,
                             "peak_source": hbm_src, "traffic": traffic.get("k_continuum+k_raytrace"), "kernel_ms": t_hbm},
            "farfield": {"what": "default mode: distant region-I wings summed as Taylor polynomials (degree 31) per pixel tile on a "
                                 "4-level tile hierarchy (64..32768 pixels): multipole moments + tile-to-tile translations for "
                                 "pairs covering the neighbourhood (k_s2m, k_m2l), direct expansion otherwise (k_far_coeffs)",
                         "k2_ms": k2_far_ms, "reference_equivalent_region_evals_all_ranks": region_evals.tolist(),
                         "far_replaced_evals": ex["far_replaced_evals"]},
            "phase_ms": {name: phase[i] for i, name in enumerate(hp.PHASES)},
            "kernel_ms": kern,
            "parity": parity,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "nu-points/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if direct is not None:
            out["roofline_direct"] = {
                "kernel": "k_lines, direct mode (every (line, depth, pixel) Voigt evaluation done explicitly, far field off)",
                "bound": "fp64", "achieved": direct["tflops"], "peak": dfma_peak, "unit": "TFLOP/s",
                "frac": direct["tflops"] / dfma_peak, "traffic": traffic.get("k_lines_direct"), "kernel_ms": direct["k2_ms"],
                "gevals_per_s": float(stats["evals"] / direct["k2_ms"] / 1e6)}
            out["farfield"].update(k2_ms_direct=direct["k2_ms"], k2_speedup=direct["k2_ms"] / k2_far_ms,
                                   max_rel_dev_vs_direct=direct["far_dev"], ms_per_step_direct=direct["ms_per_step"],
                                   value_direct_mode=N / (direct["ms_per_step"] * 1e-3))
        print(json.dumps(out), flush=True)
        if parity is not None and not parity["ok"]:
            print(f"PARITY FAILURE on the benched workload: {parity}", file=sys.stderr, flush=True)
            rc = 3
    if world > 1:
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


# ----------------------------------------------------------------------------------------------- grid sweep (config #5)
def run_grid_sweep(args, ctx, stream, rank, world, local_rank, barrier, max_over_ranks):
    """BASELINE.json configs[4]: 64 stellar models (Teff / log g / [Fe/H] grid) scattered over the ranks as independent
    replicas -- every model is one full hot-path pass over the flagship grid on one GPU; the only collective is the
    gather of the emergent spectra.  Reports whole-box spectra/s (and nu-points/s = 64 N / time)."""
    import torch
    import torch.distributed as dist

    from stardis_b200 import _lib as L
    from stardis_b200.synthetic import sweep_models

    n_models = 64
    w, cfg, desc = build_workload(args, "solar_full")
    models = sweep_models(w, n_models)                 # per model: atmosphere-dependent inputs (host, O(D) + (L, D))
    mine = list(range(rank, n_models, world))
    hp = HotPath(ctx, w, cfg, 0, 1, torch)             # every rank evaluates the WHOLE grid of its models
    N, D = hp.N, hp.D
    resident = []
    for m in mine:                                      # inputs resident in HBM before the timed region
        mm = models(m)
        resident.append(dict(T=mm["T"], n_e=mm["n_e"], n_H=mm["n_H"], desc=mm["desc"],
                             alpha=torch.from_numpy(mm["alpha_line"]).cuda(), cont=mm["continuum"]))
    spectra = torch.empty((len(mine), N), dtype=torch.float64, device="cuda")
    gathered = [torch.empty((len(range(r, n_models, world)), N), dtype=torch.float64, device="cuda") for r in range(world)]

    def sweep():
        for i, r in enumerate(resident):
            ctx.set_atmosphere(r["T"], r["n_e"], r["n_H"], hp.model_vmic)
            hp.d_lines["alpha_line"] = r["alpha"]
            hp.bf_prefix, hp.ff_coef, hp.electron, hp.tables = r["cont"]
            hp.d_spec = spectra[i]
            hp.step(gather=False)
        if world > 1:
            if all(g.shape == gathered[0].shape for g in gathered):
                dist.all_gather(gathered, spectra)
            else:
                for src in range(world):
                    buf = spectra if src == rank else gathered[src]
                    dist.broadcast(buf, src)

    from stardis_b200 import units as u
    hp.model_vmic = float(u.cgs_values_of(hp.model.microturbulence))
    sampler = ClockSampler(local_rank)
    sampler.start()
    sweep()
    barrier()
    launches0 = ctx.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record(stream)
    for _ in range(args.steps):
        sweep()
        ctx.synchronize()
    t1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(t0.elapsed_time(t1)) / args.steps
    launches = ctx.launch_count() - launches0
    if rank == 0:
        spec = spectra.cpu().numpy()
        out = {"metric": METRIC, "value": n_models * N / (ms * 1e-3), "unit": "nu-points/s", "n_gpus": world,
               "steps": args.steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": f"grid_sweep64: {n_models} stellar models (Teff x log g x [Fe/H] variations of the solar "
                                      f"MARCS structure) x [{desc}], scattered {len(mine)} per GPU as independent replicas; "
                                      f"a step = all {n_models} models", "l2": "inputs larger than L2",
                          "partition": f"replicas: models r, r + {world}, ... on rank r; spectra gathered with NCCL"},
               "spectra_per_s": n_models / (ms * 1e-3), "ms_per_model": ms / len(mine),
               "spectrum_checks": {"finite": bool(np.isfinite(spec).all()), "positive": bool((spec > 0).all()),
                                   "distinct_models": int(len({float(s[N // 2]) for s in spec}))},
               "gpu_launches": int(launches), "clocks": clocks, "cpu_baseline": None,
               "e2e": None, "roofline": None}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
