#!/usr/bin/env python
"""bench.py — hex8 element-updates/s per explicit step on N B200s (BASELINE.json metric), with the HBM
roofline of the dominant (element) kernel, the FP64-pipe figure beside it, an end-to-end number through the
C ABI with host buffers, and the reference's CPU path timed on the same box.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n EDGE] [--material neohookean|elastic]
                  [--assembly atomic|ordered] [--impl b200|reference]

N = 1 workload: the configuration the metric is quoted on — synthetic structured hex8 cube, 400^3 = 64 M
elements, Neohookean (BASELINE.json configs[2]).  N > 1 (torchrun, one rank per GPU): weak scaling, every rank
owns an EDGE^3 brick of a (EDGE*px, EDGE*py, EDGE*pz) block and the shared-node forces are summed over NVLink
peer memory inside the step (configs[3]).  A "step" is one pass of the explicit loop body over the whole mesh.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RHO, BULK, SHEAR = 7.8, 1.6e12, 0.8e12  # test/dynamics/notched_plate_native_neohookean deck constants
MATERIAL_2 = (1.333e12, 0.1379e12, 5.0)  # bulk, shear, density of test/dynamics/brick_with_fibers material_2
# DP instruction counts are NOT constants of this file: they come from nsm_b200_kernel_info(), i.e. from the SASS of
# the library that runs (scripts/sass_hot_loop.py at build time).  The ncu DRAM-traffic figure cannot be derived
# from the binary; it is read from profiles/ncu_traffic.json, which records the kernel-source hash it was captured
# at, and is reported only while that hash equals the running library's (otherwise `traffic` is null and says why).
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")


def env_int(k, d):
    return int(os.environ.get(k, d))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md).  start() returns once
    the first sample has arrived; stop() keeps the samples taken between mark_begin() and mark_end()."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t_end = time.time() + 3.0
            while not self.rows and time.time() < t_end:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0, t1 = self.t0 or 0.0, (self.t1 or time.time()) + 0.06  # a row is stamped when read, up to one period late
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 and len(r) >= 9]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(device):
    """Pin this rank's host threads (and hence its first-touch / pinned allocations) to the CPUs next to its GPU: with 8
    ranks on a two-socket box the host<->device copies of the end-to-end step otherwise cross the socket interconnect
    (round 1: 28 % end-to-end weak-scaling efficiency at N = 8 with 110 GB/s of aggregate host traffic)."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(device), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=30).stdout.strip().lower()
        dev = bus[-12:] if len(bus) >= 12 else bus  # sysfs spells the domain with four digits
        base = "/sys/bus/pci/devices/" + dev
        cpulist = open(base + "/local_cpulist").read().strip()
        node = int(open(base + "/numa_node").read().strip())
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"pci": dev, "numa_node": node, "cpus": cpulist}
    except Exception as ex:
        return {"error": repr(ex)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def initial_velocity(mesh):
    v = np.zeros((len(mesh["x"]), 3))
    v[:, 0] = 1000.0 * mesh["x"]  # SURVEY.md §8d loading: initial_velocity x "1000.0*x"
    return v


CPU_SAMPLE_EDGE = 128  # the CPU arms time a 128^3-element cube: 2.1 M elements, ~0.5 GB of state, well out of L3


def cpu_baseline(material, cores, budget_s=15.0, single_core=False, n=CPU_SAMPLE_EDGE):
    """The reference's own serial per-element code (oracle/_ref, compiled from /root/reference/src) on `cores`
    host threads, bounded sample; falls back to the plain-C port when the reference objects are absent."""
    from nimblesm_b200.mesh import structured_cube
    from oracle import hex8 as port
    from oracle import refdrive

    mesh = structured_cube(n)
    conn = mesh["conn"][1]
    ref = np.ascontiguousarray(np.stack([mesh["x"], mesh["y"], mesh["z"]], 1))
    mass = port.lumped_mass(RHO, ref, conn)
    dt = 0.2 * (1.0 / n) / np.sqrt(BULK / RHO)
    kind = "reference" if refdrive.available() else "port"

    def run(steps, threads):
        u, v, a = np.zeros_like(ref), initial_velocity(mesh), np.zeros_like(ref)
        if kind == "reference":
            ms = "%s density %r bulk_modulus %r shear_modulus %r" % (material, RHO, BULK, SHEAR)
            t, _ = refdrive.bench_steps(ms, ref, conn, mass, u, v, a, dt, steps, threads)
        else:
            t, _ = port.bench_steps(port.NEOHOOKEAN if material == "neohookean" else port.ELASTIC, BULK, SHEAR, ref,
                                    conn, mass, u, v, a, dt, steps, threads)
        return t

    t1 = run(1, cores)
    steps = int(max(2, min(400, budget_s / max(t1, 1e-3))))
    t = run(steps, cores)
    out = {"value": len(conn) * steps / t, "unit": "element-updates/s", "cores": cores, "kind": kind,
           "sample": "%d^3 hex8 cube (%d elements), %s, %d explicit steps, %d threads over element chunks, %.1f s"
                     % (n, len(conn), material, steps, cores, t)}
    if single_core:
        # the reference's own build is serial (its only multi-core path is Kokkos-OpenMP): one thread,
        # on a smaller sample so that it stays within the budget
        out["single_core_value"] = len(conn) * 1 / run(1, 1)
    return out, t / steps


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref = its serial sources compiled
    in place; the plain-C port only if those objects are missing) on all host threads.  `config` is the workload of
    the B200 arm (same dict); what is actually timed each step is the bounded sample named in `cpu_baseline.sample`
    and `sample_config`, and `ms_per_step` is that measurement extrapolated to one step of the full workload
    (elements of the workload / measured element-updates per second; the per-element cost does not depend on the
    mesh size once the sample is out of cache)."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    world = env_int("WORLD_SIZE", args.gpus)
    material = "neohookean" if args.workload == "twoblock" else args.material
    # K timed "steps", each a bounded sample of the workload
    per = max(3.0, min(20.0, 120.0 / max(args.steps + args.warmup, 1)))
    vals, sample_ms, last = [], [], None
    for i in range(args.warmup + args.steps):
        cb, sec_per_step = cpu_baseline(material, cores, budget_s=per)
        if i >= args.warmup:
            vals.append(cb["value"])
            sample_ms.append(sec_per_step * 1e3)
        last = cb
    v = float(np.mean(vals))
    last["value"] = v
    cfg = workload_config(args, args.n, world)
    print(json.dumps({
        "impl": "reference", "metric": "hex8 element-updates/sec per explicit step", "value": v,
        "unit": "element-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cfg["elements_per_gpu"] * world / v * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "sample_config": {"workload": "%d^3-element sample of config.workload per timed step (bounded CPU time); "
                                      "ms_per_step is extrapolated to the full workload" % CPU_SAMPLE_EDGE,
                          "elements": CPU_SAMPLE_EDGE ** 3, "sample_ms_per_step": float(np.mean(sample_ms))},
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": "element-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, n_edge, world=1):
    """The workload, from the command line alone: both arms print the same dict."""
    from nimblesm_b200.mesh import brick_grid

    two = getattr(args, "workload", "cube") == "twoblock"
    mat = "elastic block 1 (x < mid) + neohookean block 2, prescribed velocity on both x faces" if two else args.material
    px, py, pz = brick_grid(world)
    return {"workload": "synthetic structured hex8 cube %d^3 = %d elements per GPU, %s, explicit central difference"
                        % (n_edge, n_edge ** 3, mat),
            "elements_per_gpu": n_edge ** 3, "material": "elastic+neohookean" if two else args.material,
            "assembly": args.assembly, "numbering": "random (shuffled)" if getattr(args, "shuffle", False) else "lattice order",
            "l2_policy": "inputs larger than L2 (per-step traffic >> 126 MB)" if n_edge ** 3 * 234 > 4e8
                         else "small workload: L2-resident",
            "grid": [px, py, pz], "nodes_per_gpu": (n_edge + 1) ** 3,
            "dt": 0.2 * (1.0 / (n_edge * px)) / float(np.sqrt(BULK / RHO)),
            "cpu_reference_sample": "the reference arm (--impl reference) and cpu_baseline time a %d^3-element sample "
                                    "of this workload on the host cores" % CPU_SAMPLE_EDGE}


def dram_traffic(kernel_key, source_sha, n_elem):
    """ncu DRAM bytes per element of the element kernel, valid only for the kernel sources it was captured at."""
    try:
        rec = json.load(open(TRAFFIC_FILE))
    except Exception:
        return None, "no ncu capture on file (%s)" % os.path.relpath(TRAFFIC_FILE, ROOT)
    if rec.get("source_sha") != source_sha:
        return None, ("stale: %s was captured at kernel-source hash %s, the running library is %s; re-run "
                      "scripts/ncu_traffic.sh" % (os.path.relpath(TRAFFIC_FILE, ROOT), rec.get("source_sha"), source_sha))
    e = rec.get("kernels", {}).get(kernel_key)
    if not e:
        return None, "no capture of %s in %s" % (kernel_key, os.path.relpath(TRAFFIC_FILE, ROOT))
    return e["dram_bytes_per_element"] * n_elem, ("ncu --set full capture of this kernel build (%s, %s elements), "
                                                 "dram__bytes_read.sum + dram__bytes_write.sum per launch scaled per "
                                                 "element" % (rec.get("report", "profiles/"), e.get("elements")))


def copy_ceiling(device, world, dist, nbytes):
    """All ranks at once, pinned host memory, no compute: one-way H2D, one-way D2H, and both directions together
    (two streams).  Aggregate GB/s over the ranks = what bounds `e2e` on this box however the step is scheduled."""
    import torch

    torch.cuda.set_device(device)
    h_up, h_dn = torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_up, d_dn = torch.empty(nbytes, dtype=torch.uint8, device="cuda"), torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down, reps=3):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s_up):
                    d_up.copy_(h_up, non_blocking=True)
            if down:
                with torch.cuda.stream(s_dn):
                    h_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        return (int(up) + int(down)) * nbytes * reps * world / dt / 1e9

    run(True, True, 1)
    return {"bytes_per_copy": nbytes, "h2d_only_gbs_all_ranks": run(True, False), "d2h_only_gbs_all_ranks": run(False, True),
            "both_directions_gbs_all_ranks": run(True, True),
            "what": "pinned-memory copies by every rank at the same time, nothing else running"}


def parity_check(c, mesh, args, n, grid, pos, rank, world, dist):
    """CHECKER, after every timed region (the oracle is test infrastructure; nothing timed goes through it).
    (1) N > 1: every replica of every shared node carries bit-identical u, v, f_int on all its holders (the
        reference pins its reducer the same way: np2/np4 runs against the serial gold, test/run_exodiff_test.py:
        103-111,164-190).
    (2) A 2w^3-element window around the point where the partitions meet (the cube centre at N = 1), gathered from
        all ranks by global lattice position: with the GPUs' own displacement the CPU oracle's internal force must
        match the device's on every node the window determines, to 1e-12 (max-norm) -- across partition faces this
        checks the shared-node sum itself."""
    from nimblesm_b200.mesh import HEX_CORNERS

    px, py, pz = grid
    N = (n * px, n * py, n * pz)
    u, v, f = (c.download(k) for k in ("displacement", "velocity", "internal_force"))
    out = {"checked": True}
    # ---- (1) replicas
    if world > 1:
        surf = mesh["surface_idx"]
        blob = (mesh["node_gid"][surf], u[surf], v[surf], f[surf])
        allb = [None] * world
        dist.all_gather_object(allb, blob)
        if rank == 0:
            gid = np.concatenate([b[0] for b in allb])
            order = np.argsort(gid, kind="stable")
            gid = gid[order]
            same = gid[1:] == gid[:-1]
            ok = True
            for k in (1, 2, 3):
                val = np.concatenate([b[k] for b in allb])[order].view(np.int64)
                ok = ok and bool(np.all(val[1:][same] == val[:-1][same]))
            out["replicas_bit_equal"] = ok
            out["shared_node_replicas_compared"] = int(same.sum())
    # ---- (2) window vs oracle
    w = 8 if min(N) >= 32 else max(1, min(N) // 4)
    cen = [n if p > 1 else (n // 2) for p in (px, py, pz)]  # global element index of the meeting point
    lo = [max(0, min(cen[d] - w, N[d] - 2 * w)) for d in range(3)]
    W = 2 * w
    nn = n + 1
    off = [pos[d] * n for d in range(3)]
    # this rank's nodes inside the window's node range [lo, lo + W] per axis
    rng_ax = [np.arange(max(lo[d], off[d]), min(lo[d] + W, off[d] + n) + 1, dtype=np.int64) for d in range(3)]
    if all(len(r) for r in rng_ax):
        K, J, I = np.meshgrid(rng_ax[2], rng_ax[1], rng_ax[0], indexing="ij")
        loc = ((I - off[0]) + nn * ((J - off[1]) + nn * (K - off[2]))).ravel()
        wid = ((I - lo[0]) + (W + 1) * ((J - lo[1]) + (W + 1) * (K - lo[2]))).ravel()
        X = np.stack([mesh["x"][loc], mesh["y"][loc], mesh["z"][loc]], 1)
        part = (wid, X, u[loc], f[loc])
    else:
        part = (np.zeros(0, np.int64), np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 3)))
    parts = [part]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, part)
    if rank != 0:
        return None
    from oracle import hex8 as oracle

    nw = (W + 1) ** 3
    Xw, uw, fw, seen = np.zeros((nw, 3)), np.zeros((nw, 3)), np.zeros((nw, 3)), np.zeros(nw, bool)
    for wid, X, uu, ff in parts:
        Xw[wid], uw[wid], fw[wid] = X, uu, ff
        seen[wid] = True
    assert seen.all(), "window nodes missing from every rank"
    e = np.arange(W ** 3, dtype=np.int64)
    ei, ej, ek = e % W, (e // W) % W, e // (W * W)
    conn = np.empty((W ** 3, 8), dtype=np.int32)
    for k_, (di, dj, dk) in enumerate(HEX_CORNERS):
        conn[:, k_] = (ei + di) + (W + 1) * ((ej + dj) + (W + 1) * (ek + dk))
    if args.workload == "twoblock":  # block 1 = elastic where the brick-local element index i < n/2, else block 2
        first = ((ei + lo[0]) % n) < n // 2
        f_or = np.zeros((nw, 3))
        for sel, kind, k_, g_ in ((first, oracle.ELASTIC, BULK, SHEAR), (~first, oracle.NEOHOOKEAN, MATERIAL_2[0], MATERIAL_2[1])):
            if sel.any():
                fb, _ = oracle.internal_force(kind, k_, g_, Xw, uw, np.ascontiguousarray(conn[sel]), False)
                f_or += fb
    else:
        kind = oracle.NEOHOOKEAN if args.material == "neohookean" else oracle.ELASTIC
        f_or, _ = oracle.internal_force(kind, BULK, SHEAR, Xw, uw, conn, False)
    a = np.arange(W + 1)
    Kn, Jn, In = np.meshgrid(a, a, a, indexing="ij")
    complete = np.ones(nw, bool)
    for loc_ax, d in ((In, 0), (Jn, 1), (Kn, 2)):
        g = loc_ax.ravel() + lo[d]
        complete &= ((g > lo[d]) | (g == 0)) & ((g < lo[d] + W) | (g == N[d]))
    scale = np.abs(f_or).max()
    out["max_rel_f"] = float(np.abs(fw[complete] - f_or[complete]).max() / scale) if scale > 0 else None
    out["window"] = "%d^3 elements at global element (%d, %d, %d) of the %dx%dx%d lattice, %d nodes compared, %s" % (
        W, lo[0], lo[1], lo[2], N[0], N[1], N[2], int(complete.sum()),
        "straddling the partition faces" if world > 1 else "cube centre")
    out["force_scale"] = float(scale)
    out["ok"] = bool(out["max_rel_f"] is not None and out["max_rel_f"] <= 1e-12 and out.get("replicas_bit_equal", True))
    out["what"] = ("after the timed regions: oracle (oracle/hex8_oracle.c, pinned to the reference's compiled serial code) "
                   "internal force of the window with the GPUs' own displacement vs the device force; bar 1e-12 max-norm")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", "--edge", dest="n", type=int, default=400,
                    help="cube edge in elements per GPU (400 -> 64 M elements); spell it --edge under torchrun")
    ap.add_argument("--material", default="neohookean", choices=["neohookean", "elastic", "j2_plasticity"],
                    help="j2_plasticity: the history-dependent model of the state-variable slot (yield 5e8, hardening 2e10); "
                         "its parity is pinned by tests/test_gpu_state.py, the bench line only measures it")
    ap.add_argument("--workload", default="cube", choices=["cube", "twoblock", "contact"],
                    help="twoblock: BASELINE.json configs[4], every brick split at its mid-x plane into an elastic "
                         "block 1 and a neohookean block 2 (brick_with_fibers material_2 constants), prescribed "
                         "velocity on both global x faces; contact: two stacked bodies of EDGE x EDGE x EDGE/2 neohookean "
                         "elements, the upper one falling onto the lower one (penalty contact, SURVEY §8 f-4); one GPU")
    ap.add_argument("--assembly", default="atomic", choices=["atomic", "ordered"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--flags", type=int, default=2, help="nsm_b200_finalize flags (2 = cache the reference Jacobians)")
    ap.add_argument("--shuffle", action="store_true",
                    help="random node numbering and element order (an unstructured mesh's worst case for gather locality)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to the CPUs next to its GPU")
    ap.add_argument("--copy-ceiling", action="store_true",
                    help="also measure what the box's host<->device path sustains with all ranks copying at once (no compute): "
                         "the ceiling of the end-to-end figure")
    ap.add_argument("--host-chunks", type=int, default=-1, help="node chunks of the pipelined nsm_b200_step_host (-1 auto, 0 off)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run oracle / replica check")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "contact":
        return run_contact(args)

    from nimblesm_b200 import capi
    from nimblesm_b200.mesh import brick_surface_gids, cube_partition, brick_grid, shared_node_tables

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    numa = bind_to_gpu_numa_node(local_rank) if not args.no_numa_bind else {"disabled": True}
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- mesh: this rank's brick of the weak-scaled block ------------------------------------------
    px, py, pz = brick_grid(world)
    assert px == py == pz or world in (1, 2, 4, 8)
    # weak scaling needs EDGE^3 per rank: global lattice edge differs per axis, so build the brick by hand
    from nimblesm_b200.mesh import structured_brick  # noqa: E402

    n = args.n
    rx, ry, rz = rank % px, (rank // px) % py, rank // (px * py)
    mesh = weak_brick(n, (px, py, pz), (rx, ry, rz), args.workload == "twoblock")
    n_elem, n_nodes = sum(len(cn) for cn in mesh["conn"].values()), len(mesh["x"])
    if args.shuffle:
        assert world == 1, "--shuffle is a single-GPU experiment"
        rng = np.random.default_rng(7)
        perm = rng.permutation(n_nodes)           # new id of old node i
        inv = np.empty_like(perm)
        inv[perm] = np.arange(n_nodes)
        for k in ("x", "y", "z"):
            mesh[k] = np.ascontiguousarray(mesh[k][inv])
        for bid in list(mesh["conn"]):
            cn = perm[mesh["conn"][bid]].astype(np.int32)
            mesh["conn"][bid] = np.ascontiguousarray(cn[rng.permutation(len(cn))])
        mesh["node_sets"] = {k: np.sort(perm[v]).astype(np.int32) for k, v in mesh["node_sets"].items()}

    c = capi.Context(local_rank)
    c.set_nodes(mesh["x"], mesh["y"], mesh["z"])
    if args.workload == "twoblock":
        c.add_block(1, mesh["conn"][1], "elastic", BULK, SHEAR, RHO)
        c.add_block(2, mesh["conn"][2], "neohookean", *MATERIAL_2)
    else:
        c.add_block(1, mesh["conn"][1], args.material, BULK, SHEAR, RHO, *((5.0e8, 2.0e10) if args.material == "j2_plasticity" else ()))
    c.finalize(capi.ASSEMBLY_ORDERED if args.assembly == "ordered" else capi.ASSEMBLY_ATOMIC, args.flags)
    if world > 1:
        import torch

        cand = mesh["surface_gid"]
        sizes = [None] * world
        dist.all_gather_object(sizes, int(len(cand)))
        mx = max(sizes)
        buf = torch.full((mx,), -1, dtype=torch.int64, device="cuda")
        buf[:len(cand)] = torch.from_numpy(cand).cuda()
        allb = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(allb, buf)
        cands = [b.cpu().numpy()[:s] for b, s in zip(allb, sizes)]
        peers, offs, nodes = shared_node_tables(rank, cands, mesh["node_gid"])
        c.comm_init(rank, world, peers, offs, nodes)
        blobs = [None] * world
        dist.all_gather_object(blobs, c.comm_export())
        for p in peers:
            c.comm_attach(int(p), blobs[int(p)])
        c.comm_ready()
        dist.barrier()
    h = 1.0 / (n * px)
    dt_user = 0.2 * h / np.sqrt(BULK / RHO)
    crit = c.compute_lumped_mass()
    v0 = initial_velocity(mesh)
    c.upload("velocity", v0)
    # prescribed_velocity 0 on the x = 0 face
    face = mesh["node_sets"][2]
    far = mesh["node_sets"][3] if args.workload == "twoblock" else face[:0]  # prescribed_velocity x 1000 on x = L
    if len(face) + len(far):
        nodes = np.concatenate([np.repeat(face, 3), far]).astype(np.int32)
        comps = np.concatenate([np.tile(np.arange(3, dtype=np.int32), len(face)), np.zeros(len(far), np.int32)])
        c.set_bc_table(nodes, comps, np.zeros(len(nodes), np.int32))
        c.set_bc_values(np.concatenate([np.zeros(3 * len(face)), np.full(len(far), 1000.0)]))

    def barrier():
        c.sync()
        if dist is not None:
            dist.barrier()
            import torch

            torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------------------------
    t = c.step(args.warmup, 0.0, dt_user)
    sampler = ClockSampler(local_rank)
    sampler.start()
    c.profile(True)
    launches0 = c.launch_count
    barrier()
    sampler.mark_begin()
    c.timer_start()
    t = c.step(args.steps, t, dt_user)
    ms = c.timer_stop()
    sampler.mark_end()
    barrier()
    clocks = sampler.stop()
    launches = c.launch_count - launches0
    elem_ms, node_ms, nprof = c.profile_read()
    c.profile(False)
    if dist is not None:
        import torch

        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    total_elems = n_elem * world
    value = total_elems * args.steps / (ms * 1e-3)

    # ---- end to end through the C ABI with host buffers ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        c.set_host_step_chunks(args.host_chunks)
        pin = {k: capi.PinnedArray((n_nodes, 3)) for k in ("u", "v", "a", "f")}
        for k, fld in (("u", "displacement"), ("v", "velocity"), ("a", "acceleration")):
            c.download(fld, pin[k].array)
        k_e2e = max(3, min(args.steps, 5))

        def e2e_step(tcur):
            # nsm_b200_step_host: H2D of u, v, a; the step; D2H of u (behind the element kernel), f_int, v, a
            return c.step_host(tcur, dt_user, pin["u"].array, pin["v"].array, pin["a"].array, pin["f"].array)

        t = e2e_step(t)
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            t = e2e_step(t)
        barrier()
        dt_wall = time.perf_counter() - t0
        if dist is not None:
            import torch

            tt = torch.tensor([dt_wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt_wall = float(tt.item())
        h2d, d2h = int(3 * 24 * n_nodes), int(4 * 24 * n_nodes)
        e2e = {"value": total_elems * k_e2e / dt_wall, "unit": "element-updates/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": k_e2e, "ms_per_step": dt_wall / k_e2e * 1e3,
               # both directions are in flight for the whole call in the pipelined step, so each rate is bytes / step time
               "h2d_gbs_per_rank": h2d * k_e2e / dt_wall / 1e9, "d2h_gbs_per_rank": d2h * k_e2e / dt_wall / 1e9,
               "host_traffic_gbs_all_ranks": (h2d + d2h) * world * k_e2e / dt_wall / 1e9,
               "host_chunks": args.host_chunks, "numa_binding": numa,
               "what": "per step: nsm_b200_step_host on pinned host [n][3] views = upload u,v,a, one explicit step, download "
                       "u,v,a,f_int, pipelined over node chunks (upload, element kernels and download overlap); host wall "
                       "clock, max over ranks"}
        # the same call when the caller does not ask for the internal force (the integrator reads it on output steps only)
        if world == 1:
            t = c.step_host(t, dt_user, pin["u"].array, pin["v"].array, pin["a"].array, None)
            barrier()
            t0 = time.perf_counter()
            for _ in range(k_e2e):
                t = c.step_host(t, dt_user, pin["u"].array, pin["v"].array, pin["a"].array, None)
            barrier()
            dt_state = time.perf_counter() - t0
            e2e["without_force_download"] = {"value": total_elems * k_e2e / dt_state, "unit": "element-updates/s",
                                             "ms_per_step": dt_state / k_e2e * 1e3, "h2d_bytes_per_step": h2d,
                                             "d2h_bytes_per_step": int(3 * 24 * n_nodes),
                                             "what": "nsm_b200_step_host with internal_force = NULL (a non-output step): u, v, a up and down"}
        # second figure: the reference's own crossing pattern.  Its Kokkos path keeps the integrator's axpys on the host
        # and crosses once per step at ModelData::ComputeInternalForce (displacement up, internal force down;
        # src/nimble_kokkos_model_data.cc:1730-1740, 1282) -- here nsm_b200_internal_force_host on the same pinned views,
        # pipelined over node chunks.  The host-side axpys of that sequence are the caller's and are not timed.
        if world == 1:
            c.internal_force_host(pin["u"].array, out=pin["f"].array)
            barrier()
            t0 = time.perf_counter()
            for _ in range(k_e2e):
                c.internal_force_host(pin["u"].array, out=pin["f"].array)
            barrier()
            dt_seam = time.perf_counter() - t0
            e2e["force_seam"] = {"value": total_elems * k_e2e / dt_seam, "unit": "element-updates/s", "ms_per_call": dt_seam / k_e2e * 1e3,
                                 "h2d_bytes_per_step": int(24 * n_nodes), "d2h_bytes_per_step": int(24 * n_nodes),
                                 "what": "per step: nsm_b200_internal_force_host = ModelData::ComputeInternalForce on pinned host "
                                         "views (u up, f_int down, pipelined); the integrator's nodal updates stay with the caller, "
                                         "as in the reference's Kokkos sequence"}
        for p_ in pin.values():
            p_.free()
        if args.copy_ceiling:
            e2e["copy_ceiling"] = copy_ceiling(local_rank, world, dist, int(24 * n_nodes))

    # ---- parity of what was just timed (checker; all ranks take part, rank 0 reports) ------------------
    parity = None
    if args.material == "j2_plasticity":
        parity = {"checked": False, "ok": True, "why": "a window check needs the whole state history; the state-variable slot is "
                  "pinned bit for bit by tests/test_gpu_state.py and tests/test_gpu_ref_binding.py"}
    elif not args.no_parity:
        try:
            parity = parity_check(c, mesh, args, n, (px, py, pz), (rx, ry, rz), rank, world, dist)
        except Exception as ex:  # the checker is test infrastructure; report, do not hide
            parity = {"checked": False, "error": repr(ex)}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------------
    peaks, how = measured_peaks()
    r = n_nodes / n_elem
    elem_bytes = 32.0 + 72.0 * r  # conn 32 B + X,u read once per node 48 B + f written once per node 24 B
    step_bytes = 32.0 + 200.0 * r  # + node kernel: f 24, m 8, v 24, u 24 read, v 24, u 24 write (SURVEY §8d B_min)
    ach = elem_bytes * n_elem / (elem_ms * 1e-3) / 1e9 if elem_ms > 0 else None
    dadd, dfma = c.fp64_peak()
    # a kernel timed inside a LONG run is held to the sustained FP64 rate (power capping settles the SM clock after a few
    # hundred milliseconds); a short timed region to the burst rate
    long_run = ms > 1000.0
    dadd_sustained = c.fp64_peak_sustained(2.0) if long_run else None
    peak_used = dadd_sustained if long_run else dadd
    # FP64 work per element-update: the static DP instruction count of one warp pass of the kernel instance that
    # ran, from the library's own SASS (nsm_b200_kernel_info)
    kinfo = capi.kernel_info()
    mode = 2 if c.effective_flags & 2 else 0
    mat_id = {"elastic": 0, "neohookean": 1, "j2_plasticity": 2}
    kkey = lambda mat: "mat%d_ordered%d_mode%d" % (mat_id[mat], 1 if args.assembly == "ordered" else 0, mode | (1 if mat == "j2_plasticity" else 0))
    if args.workload == "twoblock":  # the two blocks' element kernels run back to back; figures are per-element means
        keys = [kkey("elastic"), kkey("neohookean")]
    else:
        keys = [kkey(args.material)]
    dp = float(np.mean([kinfo["kernels"][k]["dp_lane_instr_per_element"] for k in keys]))
    traffic, traffic_source = dram_traffic(keys[-1], kinfo["source_sha"], n_elem) if len(keys) == 1 else (None, "two kernels per step")
    roof = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": (ach / peaks["hbm_gbs"]) if ach else None,
            "traffic": traffic, "traffic_source": traffic_source,
            "peak_source": how,
            "kernel": "element_force_kernel", "kernel_ms": elem_ms, "kernel_share_of_step": elem_ms * nprof / ms if ms else None,
            "algorithmic_bytes_per_element": elem_bytes,
            "whole_step": {"bytes_per_element_update": step_bytes,
                           "achieved": step_bytes * n_elem * args.steps / (ms * 1e-3) / 1e9,
                           "frac": step_bytes * n_elem * args.steps / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}}
    fp64 = {"note": "the kernel is FP64-issue bound (SURVEY.md §0.5); fraction of the measured DADD/DMUL issue peak",
            "dp_instr_per_element": dp, "dp_source": "nsm_b200_kernel_info(): static SASS count of the running binary, kernel "
            + "+".join(keys) + ", sources " + kinfo["source_sha"],
            "hot_loop": {k: {x: kinfo["kernels"][k][x] for x in ("dp", "other", "hot_instructions", "reg", "stack")} for k in keys},
            "achieved_tera_lane_ops": dp * n_elem / (elem_ms * 1e-3) / 1e12 if elem_ms else None,
            "peak_dadd_dmul_tera_lane_ops": dadd, "peak_dfma_tera_lane_ops": dfma,
            "peak_dadd_dmul_sustained_tera_lane_ops": dadd_sustained,
            "peak_used": "sustained (timed region %.1f s)" % (ms * 1e-3) if long_run else "burst",
            "frac": (dp * n_elem / (elem_ms * 1e-3) / 1e12 / peak_used) if elem_ms and peak_used else None}
    out = {
        "metric": "hex8 element-updates/sec per explicit step", "value": value, "unit": "element-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n, world), "critical_dt": crit, "device_bytes": c.device_bytes,
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "fp64": fp64, "node_kernels_ms": node_ms,
        "cold_points": c.cold_points,
    }
    if e2e:
        out["e2e"] = e2e
    if parity is not None:
        out["parity"] = parity
    if not args.no_cpu and world == 1:
        try:
            out["cpu_baseline"], _ = cpu_baseline(args.material, os.cpu_count() or 1, single_core=True)
        except Exception as ex:  # the checker libraries are test infrastructure; report, do not hide
            out["cpu_baseline"] = {"error": str(ex)}
    print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    bad = []
    if not clocks.get("samples"):
        bad.append("no nvidia-smi clock sample fell inside the timed region: this line is not a valid measurement")
    if parity is not None and not parity.get("ok", False):
        bad.append("parity check failed: %r" % (parity,))
    if bad:
        sys.stderr.write("bench.py: " + "; ".join(bad) + "\n")
        sys.exit(3)


def contact_stack(n):
    """Two bodies of n x n x n/2 unit-aspect elements (h = 1/n): block 2 (primary) fills z in [0, 1/2], block 1
    (secondary) sits a thousandth of h above it, shifted by (0.37 h, 0.21 h) so that its bottom nodes meet the lower
    body's top facets at generic positions.  Returns the mesh, the contact entities as ContactManager builds them for a
    structured body (all six sides of each skin; Exodus face orders, outward normals) and the node sets."""
    from nimblesm_b200.mesh import HEX_CORNERS

    h, nz = 1.0 / n, n // 2
    face_of_side = {"z-": (4, [0, 3, 2, 1]), "z+": (5, [4, 5, 6, 7]), "x-": (3, [0, 4, 7, 3]), "x+": (1, [1, 2, 6, 5]),
                    "y-": (0, [0, 1, 5, 4]), "y+": (2, [2, 3, 7, 6])}

    def body(origin, node_base):
        nx = ny = n + 1
        idx = np.arange(nx * ny * (nz + 1), dtype=np.int64)
        i, j, k = idx % nx, (idx // nx) % ny, idx // (nx * ny)
        xyz = [origin[0] + i * h, origin[1] + j * h, origin[2] + k * h]
        e = np.arange(n * n * nz, dtype=np.int64)
        ei, ej, ek = e % n, (e // n) % n, e // (n * n)
        conn = np.empty((len(e), 8), dtype=np.int32)
        for c_, (di, dj, dk) in enumerate(HEX_CORNERS):
            conn[:, c_] = node_base + (ei + di) + nx * ((ej + dj) + ny * (ek + dk))
        sides = {"z-": ek == 0, "z+": ek == nz - 1, "x-": ei == 0, "x+": ei == n - 1, "y-": ej == 0, "y+": ej == n - 1}
        quads = np.concatenate([conn[sel][:, face_of_side[sd][1]] for sd, sel in sides.items()])
        return xyz, conn, quads, (i, j, k)

    lo_xyz, lo_conn, lo_quads, lo_ijk = body((0.0, 0.0, 0.0), 0)
    n_lo = len(lo_xyz[0])
    up_xyz, up_conn, up_quads, _ = body((0.37 * h, 0.21 * h, 0.5 + 1.0e-3 * h), n_lo)
    mesh = dict(x=np.concatenate([lo_xyz[0], up_xyz[0]]), y=np.concatenate([lo_xyz[1], up_xyz[1]]),
                z=np.concatenate([lo_xyz[2], up_xyz[2]]), block_ids=[1, 2], conn={1: up_conn, 2: lo_conn},
                node_sets={"bottom": np.flatnonzero(lo_ijk[2] == 0).astype(np.int32),
                           "upper": np.arange(n_lo, n_lo + len(up_xyz[0]), dtype=np.int32)})
    X = np.stack([mesh["x"], mesh["y"], mesh["z"]], 1)

    def longest_edge(q):  # ContactManager::CreateContactEntities: longest edge of a skin face, model configuration
        e = [np.sqrt(((X[q[:, (a + 1) % 4]] - X[q[:, a]]) ** 2).sum(1)) for a in range(4)]
        return np.max(e, axis=0)

    flat = up_quads.ravel()
    _u, first = np.unique(flat, return_index=True)
    contact_nodes = flat[np.sort(first)].astype(np.int32)
    node_len = np.zeros(len(X))
    np.maximum.at(node_len, up_quads.ravel(), np.repeat(longest_edge(up_quads), 4))
    ent = dict(primary_quads=np.ascontiguousarray(lo_quads, dtype=np.int32), primary_char_len=longest_edge(lo_quads),
               contact_nodes=contact_nodes, contact_node_char_len=node_len[contact_nodes])
    return mesh, ent, h


def run_contact(args):
    """`--workload contact`: the explicit step WITH the contact term on one GPU.  Reports the step with
    and without contact, the contact evaluation alone, the pair counters, and a parity block (contact force of a window
    of contact nodes recomputed by the oracle on the device's own displacement)."""
    from nimblesm_b200 import capi

    if env_int("WORLD_SIZE", 1) > 1 or args.gpus > 1:
        sys.stderr.write("bench.py: --workload contact is a single-GPU line (contact across partitions runs through the C++ "
                         "driver, tests/test_gpu_contact.py)\n")
        sys.exit(2)
    n = args.n if args.n != 400 else 200
    mesh, ent, h = contact_stack(n)
    n_elem, n_nodes = sum(len(c_) for c_ in mesh["conn"].values()), len(mesh["x"])
    c = capi.Context(0)
    c.set_nodes(mesh["x"], mesh["y"], mesh["z"])
    for b in (1, 2):
        c.add_block(b, mesh["conn"][b], "neohookean", BULK, SHEAR, RHO)
    c.finalize(capi.ASSEMBLY_ORDERED if args.assembly == "ordered" else capi.ASSEMBLY_ATOMIC, args.flags)
    c.compute_lumped_mass()
    dt_user = 0.2 * h / np.sqrt(BULK / RHO)
    penalty = BULK * h  # the stiffness of one element face
    bottom = mesh["node_sets"]["bottom"]
    c.set_bc_table(np.repeat(bottom, 3), np.tile(np.arange(3, dtype=np.int32), len(bottom)), np.zeros(3 * len(bottom), np.int32))
    c.set_bc_values(np.zeros(3 * len(bottom)))
    v0 = np.zeros((n_nodes, 3))
    v0[mesh["node_sets"]["upper"], 2] = -1000.0  # closes the 1e-3 h gap in the third step
    results = {}
    for mode in ("with_contact", "without_contact"):
        if mode == "with_contact":
            c.set_contact(penalty, ent["primary_quads"], ent["primary_char_len"], ent["contact_nodes"], ent["contact_node_char_len"])
        else:
            c.set_contact(0.0, np.zeros((0, 4), np.int32), np.zeros(0), np.zeros(0, np.int32), np.zeros(0))
        c.upload("displacement", np.zeros((n_nodes, 3)))
        c.upload("acceleration", np.zeros((n_nodes, 3)))
        c.upload("velocity", v0)
        t = c.step(max(args.warmup, 3), 0.0, dt_user)
        sampler = ClockSampler(0)
        sampler.start()
        c.profile(True)
        l0 = c.launch_count
        c.sync()
        sampler.mark_begin()
        c.timer_start()
        t = c.step(args.steps, t, dt_user)
        ms = c.timer_stop()
        sampler.mark_end()
        clocks = sampler.stop()
        elem_ms, node_ms, _np = c.profile_read()
        contact_ms = c.profile_read_contact()
        c.profile(False)
        results[mode] = {"ms_per_step": ms / args.steps, "element_kernels_ms": elem_ms, "node_side_ms": node_ms, "contact_ms": contact_ms,
                         "gpu_launches": int(c.launch_count - l0), "clocks": clocks}
        if mode == "with_contact":
            stats = c.contact_stats()
            c.timer_start()
            for _ in range(20):
                c.contact_force()
            eval_ms = c.timer_stop() / 20
            # parity: a window of contact nodes around the middle of the interface, facets near them, oracle on the device's u
            from oracle import contact as contact_oracle

            u = c.download("displacement")
            fc = c.download("contact_force")
            X = np.stack([mesh["x"], mesh["y"], mesh["z"]], 1)
            cn = ent["contact_nodes"]
            cur = X[cn] + u[cn]
            w = 8.0 * h
            near = (np.abs(cur[:, 0] - 0.5) < w) & (np.abs(cur[:, 1] - 0.5) < w) & (np.abs(cur[:, 2] - 0.5) < w)
            qc = (X[ent["primary_quads"]] + u[ent["primary_quads"]]).mean(1)
            qsel = (np.abs(qc[:, 0] - 0.5) < w + 3 * h) & (np.abs(qc[:, 1] - 0.5) < w + 3 * h) & (np.abs(qc[:, 2] - 0.5) < w + 3 * h)
            want = np.zeros_like(X)
            L = contact_oracle._lib()
            pairs_w = L.h8o_contact_force(penalty, n_nodes, np.ascontiguousarray(X), np.ascontiguousarray(u), int(qsel.sum()),
                                          np.ascontiguousarray(ent["primary_quads"][qsel]).reshape(-1),
                                          np.ascontiguousarray(ent["primary_char_len"][qsel]), int(near.sum()),
                                          np.ascontiguousarray(cn[near]), np.ascontiguousarray(ent["contact_node_char_len"][near]),
                                          want, None)
            scale = np.abs(fc[cn]).max()
            err = float(np.abs(fc[cn[near]] - want[cn[near]]).max() / scale) if scale > 0 else 0.0
            # size-independent property of the WHOLE surface: action = reaction, the contact forces sum to zero
            # (summed in extended precision: the facet side is all of one sign and comes first in the numbering, so a
            # double-precision running sum carries ~1e-9 of rounding of its own before the node side cancels it)
            resid = float(np.abs(fc.astype(np.longdouble).sum(0)).max() / scale) if scale > 0 else 0.0
            parity = {"checked": True, "window_contact_nodes": int(near.sum()), "window_pairs": int(pairs_w), "max_rel_fc": err,
                      "sum_of_contact_forces_over_largest": resid,
                      "ok": bool(err <= 1e-12 and pairs_w > 0 and stats["pairs"] > 0 and resid <= 1e-10),
                      "what": "contact force on the contact nodes of a 16h window of the interface, oracle/contact_oracle.c on the "
                              "device's own displacement; bar 1e-12 of the largest nodal contact force.  Whole surface: the contact "
                              "forces sum to zero (action = reaction), bar 1e-10 of the largest"}
    wc, nc = results["with_contact"], results["without_contact"]
    out = {"metric": "hex8 element-updates/sec per explicit step", "value": n_elem * 1e3 / wc["ms_per_step"], "unit": "element-updates/s",
           "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": wc["ms_per_step"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "two stacked neohookean bodies of %d x %d x %d elements each (%d elements), the upper one falling "
                                  "onto the lower one, penalty contact between all of the lower skin (%d triangles) and the upper "
                                  "skin nodes (%d)" % (n, n, n // 2, n_elem, 4 * len(ent["primary_quads"]), len(ent["contact_nodes"])),
                      "assembly": args.assembly, "penalty": penalty, "dt": dt_user},
           "gpu_launches": wc["gpu_launches"], "clocks": wc["clocks"],
           "contact": {"step_ms_with_contact": wc["ms_per_step"], "step_ms_without_contact": nc["ms_per_step"],
                       "cost_of_contact_ms_per_step": wc["ms_per_step"] - nc["ms_per_step"],
                       "contact_ms_per_step_device": wc["contact_ms"], "contact_evaluation_alone_ms": eval_ms, "launches_per_evaluation": 3,
                       "pairs_enforced": stats["pairs"], "pairs_box_tested": stats["box_tested"],
                       "active_triangles": stats["active_faces"], "active_nodes": stats["active_nodes"],
                       "node_side_ms_with": wc["node_side_ms"], "node_side_ms_without": nc["node_side_ms"]},
           "parity": parity, "without_contact": nc}
    print(json.dumps(out))
    if not wc["clocks"].get("samples") or not parity["ok"]:
        sys.stderr.write("bench.py: contact workload: invalid line (%r)\n" % ({"clocks": wc["clocks"], "parity": parity},))
        sys.exit(3)


def weak_brick(n, grid, pos, twoblock=False):
    """EDGE^3-element brick at grid position `pos` of a (n*px, n*py, n*pz) block with spacing 1/(n*px).
    twoblock: elements with local i < n/2 form block 1, the rest block 2 (every rank holds both materials)."""
    px, py, pz = grid
    rx, ry, rz = pos
    N = (n * px, n * py, n * pz)
    h = 1.0 / (n * px)
    lo = (rx * n, ry * n, rz * n)
    nn = n + 1
    idx = np.arange(nn ** 3, dtype=np.int64)
    i, j, k = idx % nn, (idx // nn) % nn, idx // (nn * nn)
    gi, gj, gk = i + lo[0], j + lo[1], k + lo[2]
    mesh = dict(x=gi * h, y=gj * h, z=gk * h, block_ids=[1], conn={}, node_sets={})
    mesh["node_gid"] = gi + (N[0] + 1) * (gj + (N[1] + 1) * gk)
    e = np.arange(n ** 3, dtype=np.int64)
    ei, ej, ek = e % n, (e // n) % n, e // (n * n)
    conn = np.empty((n ** 3, 8), dtype=np.int32)
    from nimblesm_b200.mesh import HEX_CORNERS

    for c_, (di, dj, dk) in enumerate(HEX_CORNERS):
        conn[:, c_] = (ei + di) + nn * ((ej + dj) + nn * (ek + dk))
    if twoblock:
        mesh["block_ids"] = [1, 2]
        mesh["conn"][1] = np.ascontiguousarray(conn[ei < n // 2])
        mesh["conn"][2] = np.ascontiguousarray(conn[ei >= n // 2])
    else:
        mesh["conn"][1] = conn
    allnodes = np.arange(nn ** 3, dtype=np.int32)
    mesh["node_sets"][2] = allnodes[gi == 0]
    mesh["node_sets"][3] = allnodes[gi == N[0]]
    surf = np.zeros(nn ** 3, dtype=bool)
    for loc, r_, p_ in ((i, rx, px), (j, ry, py), (k, rz, pz)):
        if r_ > 0:
            surf |= loc == 0
        if r_ < p_ - 1:
            surf |= loc == n
    mesh["surface_idx"] = np.flatnonzero(surf)
    mesh["surface_gid"] = mesh["node_gid"][surf]
    return mesh


if __name__ == "__main__":
    main()
