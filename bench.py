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
# executed DADD+DMUL+DFMA+DSETP lane-instructions per element, from the ncu source page of the profiled kernel
# (sm__inst_executed_pipe_fp64.sum x 32 / elements, profiles/r01q_dp_counts.txt): [flags & 2 == 0 (b^-1 recomputed),
# flags & 2 (b^-1 cached)]
DP_INSTR_PER_ELEMENT = {"neohookean": (10640, 8968), "elastic": (5984, 4312)}
# dram__bytes_read.sum + dram__bytes_write.sum of one element-kernel launch / elements, same captures (200^3 cube)
DRAM_TRAFFIC_PER_ELEMENT = {"neohookean": (127.4, 721.4), "elastic": (128.2, 706.2)}


def env_int(k, d):
    return int(os.environ.get(k, d))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def initial_velocity(mesh):
    v = np.zeros((len(mesh["x"]), 3))
    v[:, 0] = 1000.0 * mesh["x"]  # SURVEY.md §8d loading: initial_velocity x "1000.0*x"
    return v


def cpu_baseline(material, cores, budget_s=15.0, single_core=False):
    """The reference's own serial per-element code (oracle/_ref, compiled from /root/reference/src) on `cores`
    host threads, bounded sample; falls back to the plain-C port when the reference objects are absent."""
    from nimblesm_b200.mesh import structured_cube
    from oracle import hex8 as port
    from oracle import refdrive

    n = 64 if cores <= 16 else 96
    mesh = structured_cube(n)
    conn = mesh["conn"][1]
    ref = np.ascontiguousarray(np.stack([mesh["x"], mesh["y"], mesh["z"]], 1))
    mass = port.lumped_mass(RHO, ref, conn)
    dt = 0.2 * (1.0 / n) / np.sqrt(BULK / RHO)
    kind = "reference" if refdrive.available() else "port"

    def run(steps):
        u, v, a = np.zeros_like(ref), initial_velocity(mesh), np.zeros_like(ref)
        if kind == "reference":
            ms = "%s density %r bulk_modulus %r shear_modulus %r" % (material, RHO, BULK, SHEAR)
            t, _ = refdrive.bench_steps(ms, ref, conn, mass, u, v, a, dt, steps, cores)
        else:
            t, _ = port.bench_steps(port.NEOHOOKEAN if material == "neohookean" else port.ELASTIC, BULK, SHEAR, ref,
                                    conn, mass, u, v, a, dt, steps, cores)
        return t

    t1 = run(1)
    steps = int(max(2, min(400, budget_s / max(t1, 1e-3))))
    t = run(steps)
    out = {"value": len(conn) * steps / t, "unit": "element-updates/s", "cores": cores, "kind": kind,
           "sample": "%d^3 hex8 cube (%d elements), %s, %d explicit steps, %d threads over element chunks, %.1f s"
                     % (n, len(conn), material, steps, cores, t)}
    if single_core:
        # the reference's own build is serial (its only multi-core path is Kokkos-OpenMP): one thread
        cores_all, cores = cores, 1
        ts = run(2)
        out["single_core_value"] = len(conn) * 2 / ts
        cores = cores_all
    return out, t / steps


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # K timed "steps", each a bounded sample of the workload
    per = max(3.0, min(20.0, 120.0 / max(args.steps + args.warmup, 1)))
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        cb, sec_per_step = cpu_baseline(args.material, cores, budget_s=per)
        if i >= args.warmup:
            vals.append(cb["value"])
        last = cb
    v = float(np.mean(vals))
    last["value"] = v
    n_edge = args.n
    print(json.dumps({
        "impl": "reference", "metric": "hex8 element-updates/sec per explicit step", "value": v,
        "unit": "element-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, n_edge),
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": "element-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, n_edge):
    two = getattr(args, "workload", "cube") == "twoblock"
    mat = "elastic block 1 (x < mid) + neohookean block 2, prescribed velocity on both x faces" if two else args.material
    return {"workload": "synthetic structured hex8 cube %d^3 = %d elements per GPU, %s, explicit central difference"
                        % (n_edge, n_edge ** 3, mat),
            "elements_per_gpu": n_edge ** 3, "material": "elastic+neohookean" if two else args.material,
            "assembly": args.assembly, "numbering": "random (shuffled)" if getattr(args, "shuffle", False) else "lattice order",
            "l2_policy": "inputs larger than L2 (per-step traffic >> 126 MB)" if n_edge ** 3 * 234 > 4e8
                         else "small workload: L2-resident"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", "--edge", dest="n", type=int, default=400,
                    help="cube edge in elements per GPU (400 -> 64 M elements); spell it --edge under torchrun")
    ap.add_argument("--material", default="neohookean", choices=["neohookean", "elastic"])
    ap.add_argument("--workload", default="cube", choices=["cube", "twoblock"],
                    help="twoblock: BASELINE.json configs[4], every brick split at its mid-x plane into an elastic "
                         "block 1 and a neohookean block 2 (brick_with_fibers material_2 constants), prescribed "
                         "velocity on both global x faces")
    ap.add_argument("--assembly", default="atomic", choices=["atomic", "ordered"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--flags", type=int, default=2, help="nsm_b200_finalize flags (2 = cache the reference Jacobians)")
    ap.add_argument("--shuffle", action="store_true",
                    help="random node numbering and element order (an unstructured mesh's worst case for gather locality)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    from nimblesm_b200 import capi
    from nimblesm_b200.mesh import brick_surface_gids, cube_partition, brick_grid, shared_node_tables

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
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
        c.add_block(1, mesh["conn"][1], args.material, BULK, SHEAR, RHO)
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
    c.profile(True)
    launches0 = c.launch_count
    barrier()
    sampler.start()
    c.timer_start()
    t = c.step(args.steps, t, dt_user)
    ms = c.timer_stop()
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
        e2e = {"value": total_elems * k_e2e / dt_wall, "unit": "element-updates/s",
               "h2d_bytes_per_step": int(3 * 24 * n_nodes), "d2h_bytes_per_step": int(4 * 24 * n_nodes),
               "steps": k_e2e, "what": "per step: nsm_b200_step_host on pinned host [n][3] views = upload u,v,a, one explicit step, "
                                       "download u,v,a,f_int (u overlaps the element kernel); host wall clock, max over ranks"}
        for p_ in pin.values():
            p_.free()

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
    fi = 1 if args.flags & 2 else 0
    if args.workload == "twoblock":  # the two blocks' element kernels run back to back; figures are per-element means
        dp = 0.5 * (DP_INSTR_PER_ELEMENT["elastic"][fi] + DP_INSTR_PER_ELEMENT["neohookean"][fi])
    else:
        dp = DP_INSTR_PER_ELEMENT[args.material][fi]
    roof = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": (ach / peaks["hbm_gbs"]) if ach else None,
            "traffic": DRAM_TRAFFIC_PER_ELEMENT[args.material][fi] * n_elem,
            "traffic_source": "ncu --set full capture of the same kernel on a 200^3 cube (profiles/), scaled per element; "
                              "flags & 2 adds the cached inverse reference Jacobians (576 B/element) to the 105 B/element "
                              "of algorithmic traffic",
            "peak_source": how,
            "kernel": "element_force_kernel", "kernel_ms": elem_ms, "kernel_share_of_step": elem_ms * nprof / ms if ms else None,
            "algorithmic_bytes_per_element": elem_bytes,
            "whole_step": {"bytes_per_element_update": step_bytes,
                           "achieved": step_bytes * n_elem * args.steps / (ms * 1e-3) / 1e9,
                           "frac": step_bytes * n_elem * args.steps / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}}
    fp64 = {"note": "the kernel is FP64-issue bound (SURVEY.md §0.5); fraction of the measured DADD/DMUL issue peak",
            "dp_instr_per_element": dp, "achieved_tera_lane_ops": dp * n_elem / (elem_ms * 1e-3) / 1e12 if elem_ms else None,
            "peak_dadd_dmul_tera_lane_ops": dadd, "peak_dfma_tera_lane_ops": dfma,
            "frac": (dp * n_elem / (elem_ms * 1e-3) / 1e12 / dadd) if elem_ms and dadd else None}
    out = {
        "metric": "hex8 element-updates/sec per explicit step", "value": value, "unit": "element-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, n), grid=[px, py, pz], nodes_per_gpu=n_nodes, dt=dt_user,
                       critical_dt=crit, device_bytes=c.device_bytes),
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "fp64": fp64, "node_kernels_ms": node_ms,
        "cold_points": c.cold_points,
    }
    if e2e:
        out["e2e"] = e2e
    if not args.no_cpu and world == 1:
        try:
            out["cpu_baseline"], _ = cpu_baseline(args.material, os.cpu_count() or 1, single_core=True)
        except Exception as ex:  # the checker libraries are test infrastructure; report, do not hide
            out["cpu_baseline"] = {"error": str(ex)}
    print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


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
    mesh["surface_gid"] = mesh["node_gid"][surf]
    return mesh


if __name__ == "__main__":
    main()
