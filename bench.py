#!/usr/bin/env python
"""bench.py -- rendered RGB-D frames/s of the per-step splat render (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement (oracle port), host cores

Workload = BASELINE.json configs[1]: the pusht gym env, 256 parallel envs x 96x96, one fixed camera,
per-env masked link poses (7 xarm6 links with the reference's real masks + the T-block segment),
robots-scene-v2 statistics (N = 118 337 canonical Gaussians, SH degree 3), synthetic parameters
(the trained checkpoints are git-LFS stubs).  A step = set the step's link poses + render every
(env, camera) image to RGB + depth + alpha.  Multi-GPU: envs are sharded, scene replicated, no
data-path collective ("weak": every rank renders 256 envs).

The same JSON line also carries: `other_configs` (configs[0], [2], [3]-shard, [4]-chunk: ms/step,
frames/s, stage times), `strong_config3` (BASELINE configs[3]: 2 048 envs x 2 cameras x 224x224 divided over the N
ranks -- the strong-scaling workload BASELINE.md section 4 names), `cpu_baseline_config0` (the oracle on one 640x480
frame, median of 5), `e2e_device` (observations handed over as CUDA tensors) and an issue-slot roofline for the
compositor next to the HBM one.

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "rendered RGB-D frames/sec (batched envs)"
UNIT = "frames/s"
E_PER_GPU, W, H, SH_DEG = 256, 96, 96, 3
WORKLOAD = ("configs[1]: pusht gym env, 256 envs x 96x96 RGB-D, 1 fixed camera, per-env masked link poses "
            "(7 xarm6-1 links + T-block), robots-scene-v2 N=118337 SH3, synthetic parameters")
CAP_ALL, CAP_SRC = 30_000_000, 8_000_000   # merged ids / sorted bucket entries (measured 18.7 M / 4.8 M on this workload)


def build_workload(rank: int):
    from sim_a_splat_b200 import scene as S
    from sim_a_splat_b200 import workloads as WL

    sc, vms, K, _, _, _, _ = WL.config1_pusht(E_PER_GPU)
    # a few distinct pose batches so consecutive steps do not repeat inputs; env seeds are rank-disjoint
    pose_sets = [S.random_link_poses(E_PER_GPU, sc.L, seed=1000 + 100_000 * s + 1_000_000 * rank) for s in range(4)]
    return sc, vms[:1], K, pose_sets


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line), through NVML
    (same counters nvidia-smi prints; a subprocess per sample is too slow for a sub-second region)."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, device: int):
        self.device, self.sm, self.mask, self._stop, self._t, self.max_mhz, self.power = device, [], 0, threading.Event(), None, None, []
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = None
            try:
                import torch

                uuid = "GPU-" + str(torch.cuda.get_device_properties(device).uuid)
            except Exception:
                pass
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode()) if uuid else pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)

    def summary(self) -> dict:
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(k for k, bit in self.REASONS.items() if self.mask & bit), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None, "source": "NVML, 5 ms period, sampled inside the timed region"}


def algorithmic_bytes(st: dict, N: int, K_sh: int, out_bytes_per_px: int) -> dict:
    """ALGORITHMIC bytes per launch (DESIGN.md 'Kernels and rooflines'): the bytes each stage must move if
    every operand of the work it actually does were read / written exactly once; counted from the
    measured counters.  Background rows are views (shared by all images with that camera), link rows
    are images -- the sharing lowers the work, and only the work done is counted."""
    I, C, nbg = st["n_images"], st["n_views"], st["n_background"]
    n_rows_g = C * nbg + I * (N - nbg)          # (row, gaussian) projections
    T = st["tile_w"] * st["tile_h"]
    P = st["width"] * st["height"] * I
    M, M_src, n_vis, staged = st["n_isects"], st["n_isects_src"], st["n_visible_src"], st["n_staged"]
    return {
        "project": n_rows_g * (48 + 12 * K_sh) + n_rows_g * 8 + n_vis * 48,
        "scan": ((C + I) * T + I * T) * 8,
        "emit": n_rows_g * 8 + M_src * 8,
        "sort": M_src * 16 + M * 12,                 # bucket sort (8 B in, 8 B out) + merge (8 B in, 4 B out)
        "composite": staged * 52 + I * T * 8 + P * out_bytes_per_px,
    }


def time_workload(torch, name, wl, steps, device, max_images=None, max_isects=0, max_isects_src=0, radius_rule="torch"):
    """ms/step, frames/s and per-stage times of one workload (device-resident poses, one step at a time)."""
    from sim_a_splat_b200.renderer import SplatRenderer

    sc, vms, K, w, h, poses, env = wl
    I = vms.shape[0]
    E = 1 if poses is None else poses.shape[0]
    r = SplatRenderer(sc, device=device, max_images=max_images or I, max_envs=E, max_width=w, max_height=h, max_isects=max_isects,
                      max_isects_src=max_isects_src, radius_rule=radius_rule)
    pd = torch.from_numpy(poses).to(f"cuda:{device}") if sc.L else None
    if sc.L:
        r.set_link_poses(pd)
    r.set_cameras(vms, K, w, h, env)
    names = ("rgb", "depth", "alpha")
    out = {n: r._alloc(n, False) for n in names}
    r.render(names, out=out, check=True)   # sizes the intersection buffers
    chunked = r._chunk is not None
    for _ in range(3):
        r.render(names, out=out, check=chunked)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        if sc.L:
            r.set_link_poses(pd)
        r.render(names, out=out, check=chunked)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    line = {"config": name, "images": I, "width": w, "height": h, "n_gaussians": sc.N, "ms_per_step": round(ms, 4),
            "frames_per_s": round(I / ms * 1e3, 1)}
    if not chunked:
        r.set_stage_timing(True)
        r.render(names, out=out, check=False)
        line["stages_ms"] = {k: round(v, 4) for k, v in r.stage_ms().items()}
        r.set_stage_timing(False)
        st = r.stats()
        line.update({k: st[k] for k in ("n_views", "n_isects", "n_isects_src", "max_tile_list")})
    else:
        line["chunk_images"] = int(r._chunk)
    r.close()
    return line


def other_configs(torch, device, steps):
    """BASELINE.json configs[0], [2], the configs[3] shard one of 8 GPUs renders, and an 8-camera chunk of configs[4]."""
    from sim_a_splat_b200 import workloads as WL

    res = []
    res.append(time_workload(torch, "configs[0] tblock pusht scene, one 640x480 frame", WL.config0_single_frame(), steps, device))
    res.append(time_workload(torch, "configs[2] xarm, 4 cameras x 1280x720, per-link updates", WL.config2_xarm(), steps, device))
    res.append(time_workload(torch, "configs[3] divar113vhw shard of one of 8 GPUs: 256 envs x 2 cameras x 224x224",
                             WL.config3_divar(2048, 0, 256), steps, device, max_isects=120_000_000, max_isects_src=40_000_000))
    res.append(time_workload(torch, "configs[4] synthetic 6 M Gaussians SH-3, 1920x1080, 8 of the 64 cameras per call",
                             WL.config4_stress(8), max(steps // 4, 3), device))
    # the bench workload under gsplat 1.5's CUDA radius rule (opacity-aware extend; identical images, fewer intersections):
    # reported for information, the headline uses the pinned default rule
    res.append(time_workload(torch, "configs[1] with radius_rule=cuda_1_5 (one step at a time; compare with 'serial')", WL.config1_pusht(E_PER_GPU),
                             steps * 2, device, max_isects=CAP_ALL, max_isects_src=CAP_SRC, radius_rule="cuda_1_5"))
    return res


def strong_config3(torch, dist, world, rank, device, steps):
    """BASELINE configs[3] as a STRONG-scaling job: the same 2 048 envs x 2 cameras x 224x224 (4 096 images) whatever
    N; rank r renders envs [r E/N, (r+1) E/N) in library calls of at most 512 images.  Time = max over ranks."""
    from sim_a_splat_b200 import dist as D
    from sim_a_splat_b200 import workloads as WL
    from sim_a_splat_b200.renderer import SplatRenderer

    E_total = 2048
    envs = D.shard_envs(E_total, world, rank)
    sc, vms, K, w, h, poses, env = WL.config3_divar(E_total, envs.start, envs.stop)
    r = SplatRenderer(sc, device=device, max_images=512, max_envs=len(envs), max_width=w, max_height=h, max_isects=120_000_000,
                      max_isects_src=40_000_000)
    pd = torch.from_numpy(poses).to(f"cuda:{device}")
    r.set_link_poses(pd)
    r.set_cameras(vms, K, w, h, env)
    names = ("rgb", "depth", "alpha")
    out = {n: r._alloc(n, False) for n in names}
    r.render(names, out=out, check=True)        # sizes the buffers (and settles the chunk size)
    for _ in range(2):
        r.render(names, out=out, check=False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r.set_link_poses(pd)
        r.render(names, out=out, check=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    assert r.poll_overflow() is None
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{device}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    chunk = int(r._chunk) if r._chunk is not None else 2 * len(envs)
    r.close()
    return {"workload": "configs[3]: divar113vhw, 2048 envs x 2 cameras x 224x224 = 4096 images in total, envs sharded over the ranks",
            "scaling": "strong", "n_gpus": world, "images_total": 2 * E_total, "images_per_rank": 2 * len(envs),
            "ms_per_step": round(ms, 4), "value": round(2 * E_total / ms * 1e3, 1), "unit": UNIT,
            "calls_per_step": -(-2 * len(envs) // chunk), "steps": steps, "timing": "CUDA events around the K steps, max over ranks"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from sim_a_splat_b200.renderer import PipelinedSplatRenderer, SplatRenderer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    numa_cpus = None
    if world > 1 and os.environ.get("SAS_BIND_NUMA", "1") != "0":
        from sim_a_splat_b200.dist import bind_to_gpu_numa

        numa_cpus = bind_to_gpu_numa(local)  # before any pinned allocation: observation copies stay on the GPU's socket

    sc, cam, K, pose_sets = build_workload(rank)
    I = E_PER_GPU
    kw = dict(device=local, max_images=I, max_envs=E_PER_GPU, max_width=W, max_height=H, max_isects=CAP_ALL, max_isects_src=CAP_SRC)
    r = SplatRenderer(sc, **kw)
    r.set_cameras(np.repeat(cam, I, axis=0), K, W, H, np.arange(I))
    poses_dev = [torch.from_numpy(p).to(dev) for p in pose_sets]
    poses_pin = [torch.from_numpy(p).pin_memory() for p in pose_sets]
    names = ("rgb", "depth", "alpha")
    out_dev = {n: r._alloc(n, False) for n in names}
    stream = torch.cuda.current_stream()

    def step_device(i):
        r.set_link_poses(poses_dev[i % len(poses_dev)])
        r.render(names, out=out_dev, check=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also checks capacity once) ----
    for i in range(max(args.warmup, 3)):
        step_device(i)
    st = r.stats()
    assert st["overflow"] == 0, st
    launches_per_step = r.last_launch_count()

    # two steps in flight on two libsas handles / streams (PipelinedSplatRenderer): `value` and the pipelined `e2e`
    pr = PipelinedSplatRenderer(sc, depth=2, **kw)
    pr.set_cameras(np.repeat(cam, I, axis=0), K, W, H, np.arange(I))

    def step_pipelined(i):
        pr.set_link_poses(poses_dev[i % len(poses_dev)])
        pr.render(names)

    for i in range(max(args.warmup, 3) + 1):
        step_pipelined(i)
    pr.join()

    # ---- timed region: inputs resident in HBM ----
    barrier()
    with ClockSampler(local) as clocks:
        # (1) `value`: K steps, two in flight -- step k+1's projection / binning / sort fill the SMs that step k's
        # compositing tail leaves idle; every step does all of its own work into its own output tensors
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for i in range(args.steps):
            step_pipelined(i)
        pr.join()
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        # (2) the same K steps strictly one after the other on one stream (step latency)
        ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev4.record(stream)
        for i in range(args.steps):
            step_device(i)
        ev5.record(stream)
        barrier()
        ms_serial = ev4.elapsed_time(ev5)
        # ---- end-to-end through the public API with HOST buffers (pinned): H2D poses + render + D2H obs ----
        e2e_names = ("rgb_u8", "depth")
        # (a) blocking: every step waits for its own observations before the next one is issued
        r.set_link_poses(poses_pin[0].numpy()); r.render(e2e_names, to_host=True, check=False)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            r.set_link_poses(poses_pin[i % len(poses_pin)].numpy())
            obs = r.render(e2e_names, to_host=True, check=False)  # ONE D2H copy + stream sync inside
        ms_e2e_sync = (time.perf_counter() - t0) * 1e3   # the last render() returned after its D2H: device work is complete

        # (b) pipelined, two steps in flight (PipelinedSplatRenderer: two libsas handles on two streams, used
        # alternately): step k+1's projection / binning / sort overlap step k's compositing tail and step k-1's D2H;
        # every step's poses are still copied H2D and every step's observations still land in host memory inside
        # the timed region, and every step does all of its own device work
        def e2e_pipelined(out_names):
            for i in range(4):
                pr.set_link_poses(poses_pin[i % len(poses_pin)].numpy()); pr.wait(pr.render_async(out_names))
            barrier()
            t0 = time.perf_counter()
            tickets = []
            o = None
            issue_s = 0.0   # host time spent issuing (pose upload + render_async), i.e. not waiting for observations
            for i in range(args.steps):
                ti = time.perf_counter()
                pr.set_link_poses(poses_pin[i % len(poses_pin)].numpy())
                tickets.append(pr.render_async(out_names))
                issue_s += time.perf_counter() - ti
                if i >= 2:
                    o = pr.wait(tickets[i - 2])  # consume step i-2 while i-1 and i are on the device: a handle's next render
                                                 # is already queued behind its current one when that one's copy is waited for
            for tk in tickets[-2:]:
                o = pr.wait(tk)
            dt = (time.perf_counter() - t0) * 1e3
            barrier()
            return dt, o, issue_s * 1e3 / args.steps

        ms_e2e, obs, ms_issue = e2e_pipelined(e2e_names)
        ms_e2e_f16, obs16, _ = e2e_pipelined(("rgb_u8", "depth_f16"))   # 5 B/px instead of 7
        # (c) observations handed over as CUDA tensors (an on-device policy): H2D poses every step, no D2H
        for i in range(3):
            pr.set_link_poses(poses_pin[i % len(poses_pin)].numpy()); pr.render(e2e_names)
        pr.join(); barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            pr.set_link_poses(poses_pin[i % len(poses_pin)].numpy())
            pr.render(e2e_names)
        pr.join()
        torch.cuda.current_stream().synchronize()
        ms_e2e_dev = (time.perf_counter() - t0) * 1e3
        barrier()
        pr.poll_overflow()          # raises if any of the unchecked renders above was incomplete
        assert all(x.stats()["overflow"] == 0 for x in pr.renderers)
        pr.close()
    assert obs["rgb_u8"].shape == (I, H, W, 3) and obs16["depth_f16"].dtype == torch.float16
    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_e2e_sync, ms_serial, ms_e2e_f16, ms_e2e_dev], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_e2e_sync, ms_serial, ms_e2e_f16, ms_e2e_dev = (float(v) for v in t)

    # ---- per-stage durations (CUDA events on the launching stream), same steps ----
    r.set_stage_timing(True)
    stage = {k: 0.0 for k in r.STAGES}
    for i in range(args.steps):
        step_device(i)
        for k, v in r.stage_ms().items():
            stage[k] += v / args.steps
    r.set_stage_timing(False)
    st = r.stats()

    # ---- optional observation gather to rank 0 (not part of the render path; reported separately) ----
    gather = None
    if world > 1:
        obs_dev = r.render(("rgb_u8",), check=False)["rgb_u8"]
        bufs = [torch.empty_like(obs_dev) for _ in range(world)] if rank == 0 else None
        for _ in range(3):
            dist.gather(obs_dev, bufs, dst=0)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(5):
            dist.gather(obs_dev, bufs, dst=0)
        g1.record(stream)
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1) / 5], device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather = {"ms": float(tg[0]), "bytes_into_rank0": int(obs_dev.numel() * (world - 1)), "what": "uint8 RGB observations, dist.gather (NCCL)"}
    r.close()

    # ---- BASELINE configs[3] as the strong-scaling workload; the other configs at N = 1 ----
    strong = None if args.no_other_configs else strong_config3(torch, dist, world, rank, local, max(args.steps // 20, 3))
    others = other_configs(torch, local, max(args.steps // 10, 5)) if (world == 1 and not args.no_other_configs) else None

    if rank == 0:
        frames = I * world * args.steps
        K_sh = (SH_DEG + 1) ** 2
        ab = algorithmic_bytes(st, sc.N, K_sh, 20)
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        nk = ROOT / "profiles" / "ncu_kernels.json"
        ncu = json.loads(nk.read_text()) if nk.exists() else {}
        dom = max((k for k in ab), key=lambda k: stage[k])
        stages = {k: {"ms": round(stage[k], 4), "alg_bytes": int(ab[k]) if k in ab else None,
                      "alg_gbs": round(ab[k] / (stage[k] * 1e-3) / 1e9, 1) if k in ab and stage[k] > 0 else None,
                      # DRAM bytes of the stage's kernels from ncu (--set full capture of this command, per launch): the
                      # algorithmic GB/s above is mostly served by the 126 MB L2, this is what reached HBM
                      "dram_bytes_ncu": ncu.get(k, {}).get("dram_bytes"),
                      "dram_gbs": round(ncu[k]["dram_bytes"] / (stage[k] * 1e-3) / 1e9, 1) if k in ncu and ncu[k].get("dram_bytes") and stage[k] > 0 else None}
                  for k in stage}
        clk = clocks.summary()
        n_sms = torch.cuda.get_device_properties(local).multi_processor_count
        sm_hz = (clk["sm_mhz"] or clk["sm_max_mhz"] or 1965.0) * 1e6
        issue_peak = n_sms * 4 * sm_hz           # one warp-instruction per scheduler per clock, 4 schedulers per SM
        cn = ncu.get(dom, {})
        t_dom = stage[dom] * 1e-3
        roof = {"kernel": dom, "bound": "issue" if dom == "composite" else "hbm"}
        hbm_part = {"achieved": round(ab[dom] / t_dom / 1e9, 1), "peak": hbm, "unit": "GB/s", "frac": round(ab[dom] / t_dom / 1e9 / hbm, 4),
                    "alg_bytes": int(ab[dom]), "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"}
        if dom == "composite" and cn.get("inst_executed"):
            ach = cn["inst_executed"] / t_dom
            roof.update({"achieved": round(ach / 1e9, 1), "peak": round(issue_peak / 1e9, 1), "unit": "G warp-instructions/s",
                         "frac": round(ach / issue_peak, 4),
                         "what": "warp-instructions per launch (ncu, profiles/ncu_kernels.json) / live CUDA-event duration, against SMs x 4 "
                                 "schedulers x the SM clock sampled in the timed region: the compositor is bound by instruction issue and "
                                 "the latency of its blend chain, not by HBM (DESIGN.md section 5)",
                         "lanes_per_instruction": cn.get("lanes_per_inst"), "fma_pipe_pct": cn.get("fma_pipe_pct"),
                         "issue_active_pct_ncu": cn.get("issue_active_pct"), "barrier_stall_share_pct": cn.get("barrier_stall_share_pct"),
                         "hbm": hbm_part})
        else:
            roof.update(hbm_part)
        roof["traffic"] = cn.get("dram_bytes")
        roof["traffic_source"] = "dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture of this command (profiles/ncu_kernels.json)"
        pairs = cn.get("evaluated_pairs")
        line = {
            "metric": METRIC, "value": frames / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": E_PER_GPU, "images_per_gpu": I, "width": W, "height": H,
                       "n_gaussians": sc.N, "sh_degree": SH_DEG, "n_links": sc.L, "parallelism": f"env-shard x{world}, scene replicated",
                       "pipeline": "two steps in flight per GPU on two libsas handles / CUDA streams (PipelinedSplatRenderer), i.e. scene and "
                                   "work buffers twice per GPU; 'serial' is the same K steps one after the other on one handle",
                       "l2": "per-step working set (records + bins + pairs + merged ids + images ~0.3 GB) exceeds the 126 MB L2 and 4 pose batches rotate; no explicit flush"},
            "serial": {"value": frames / (ms_serial * 1e-3), "unit": UNIT, "ms_per_step": ms_serial / args.steps},
            "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(pose_sets[0].nbytes), "d2h_bytes_per_step": int(I * H * W * (3 + 4)),
                    "outputs": "uint8 RGB [I,H,W,3] + f32 depth [I,H,W] into ONE pinned host block, one cudaMemcpyAsync per step",
                    "host_issue_ms_per_step": round(ms_issue, 4),   # rank 0's host time inside set_link_poses + render_async (the rest of a step it waits)
                    "mode": "PipelinedSplatRenderer.render_async / wait (two libsas handles on two streams, sas_render mem=HOST_ASYNC + "
                            "sas_wait_outputs): two steps on the device, the host consumes step k-2 after issuing step k; step k+1's "
                            "projection/sort overlaps step k's compositing tail and step k-1's D2H; host wall clock, max over ranks",
                    "blocking": {"value": frames / (ms_e2e_sync * 1e-3), "ms_per_step": ms_e2e_sync / args.steps,
                                 "mode": "SplatRenderer.render(to_host=True): each step waits for its own D2H"},
                    "depth_f16": {"value": frames / (ms_e2e_f16 * 1e-3), "ms_per_step": ms_e2e_f16 / args.steps,
                                  "d2h_bytes_per_step": int(I * H * W * (3 + 2)), "outputs": "uint8 RGB + fp16 depth (5 B/px)"}},
            "e2e_device": {"value": frames / (ms_e2e_dev * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_dev / args.steps,
                           "h2d_bytes_per_step": int(pose_sets[0].nbytes), "d2h_bytes_per_step": 0,
                           "mode": "pinned poses H2D every step, observations stay on the device as CUDA tensors (SAS_MEM_DEVICE) for an "
                                   "on-device policy; two steps in flight; host wall clock incl. the final synchronisation"},
            "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step),
            "roofline": roof,
            "roofline_step": {"alg_bytes": int(sum(ab.values())), "achieved": round(sum(ab.values()) / (sum(stage.values()) * 1e-3) / 1e9, 1),
                              "peak": hbm, "unit": "GB/s", "frac": round(sum(ab.values()) / (sum(stage.values()) * 1e-3) / 1e9 / hbm, 4)},
            "stages": stages,
            "fp32": {"staged_gaussians_per_step": st["n_staged"],
                     "evaluated_pixel_gaussian_pairs_per_launch": pairs,
                     "evaluated_pairs_per_s": round(pairs / t_dom, 1) if (pairs and dom == "composite") else None,
                     "note": "pairs = (pixel, Gaussian) sigma/alpha evaluations the compositor actually executes per launch (ncu per-instruction "
                             "thread counts of this workload, profiles/ncu_kernels.json); ~25 FLOP + one exp each"},
            "counts": {k: st[k] for k in ("n_views", "n_images", "n_background", "n_visible_src", "n_isects_src", "n_isects", "max_tile_list")},
            "vs_reference_note": "the CPU arm renders every env from scratch (the reference's semantics); this arm projects / sorts the "
                                 "background once per shared view -- without that sharing the same kernels ran 12.9 ms/step = 19.9 k frames/s "
                                 "(profiles/r01_v1_bench.json), i.e. ~180x of the ratio is hardware + kernels and ~12x is the sharing",
            "clocks": clk,
        }
        if strong:
            line["strong_config3"] = strong
        if others:
            line["other_configs"] = others
        if gather:
            line["obs_gather"] = gather
            line["cpu_affinity"] = f"rank 0 bound to the {len(numa_cpus)} CPUs NVML reports local to its GPU" if numa_cpus else "unchanged"
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(sc, cam, K, pose_sets[0], envs=32, reps=3)
            line["cpu_baseline_config0"] = cpu_baseline_config0()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_baseline(sc, cam, K, poses, envs: int, reps: int, threads: int = 0) -> dict:
    """The CPU oracle (this repo's C restatement, OpenMP over images/tiles) on a bounded sample of the same
    workload: `envs` of the 256 envs per pass.  kind = "port": the reference's own CPU path (gsplat via
    nerfstudio) is not installable here."""
    import oracle

    oracle.set_num_threads(threads or _host_threads())
    d = sc.as_oracle_dict()
    vms = np.repeat(cam, envs, axis=0)
    Ks = np.repeat(K[None], envs, axis=0)
    ts = []
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        oracle.render(d, vms, Ks, W, H, poses[:envs], np.arange(envs), keep={"rgb", "depth", "alpha"})
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts[1:]))
    return {"value": envs / t, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
            "sample": f"{envs} of the 256 envs per pass (96x96, same scene/poses/camera), median of {reps} passes after 1 warm-up; "
                      f"oracle/splat_oracle.c, OpenMP"}


def cpu_baseline_config0() -> dict:
    """SURVEY 8d / BASELINE.md section 3: configs[0] timed in full on the host cores -- one 640x480 RGB-D frame of the
    pusht scene through the oracle, median of 5 after one warm-up."""
    import oracle
    from sim_a_splat_b200 import workloads as WL

    oracle.set_num_threads(_host_threads())
    sc, vms, K, w, h, poses, env = WL.config0_single_frame()
    d = sc.as_oracle_dict()
    ts = []
    for _ in range(6):
        t0 = time.perf_counter()
        oracle.render(d, vms, K[None], w, h, poses, env, keep={"rgb", "depth", "alpha"})
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts[1:]))
    return {"value": 1.0 / t, "unit": UNIT, "ms_per_frame": t * 1e3, "cores": oracle.num_threads(), "kind": "port",
            "sample": "configs[0]: one 640x480 RGB-D frame, N=118337 SH3, median of 5 after 1 warm-up; oracle/splat_oracle.c, OpenMP"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc, cam, K, pose_sets = build_workload(0)
    import oracle

    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host thread it can
    oracle.set_num_threads(_host_threads())
    envs = 32
    d = sc.as_oracle_dict()
    vms, Ks = np.repeat(cam, envs, axis=0), np.repeat(K[None], envs, axis=0)
    for i in range(max(args.warmup, 1)):
        oracle.render(d, vms, Ks, W, H, pose_sets[i % 4][:envs], np.arange(envs), keep={"rgb", "depth", "alpha"})
    t0 = time.perf_counter()
    for i in range(args.steps):
        oracle.render(d, vms, Ks, W, H, pose_sets[i % 4][:envs], np.arange(envs), keep={"rgb", "depth", "alpha"})
    dt = time.perf_counter() - t0
    fps = envs * args.steps / dt
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"each step renders {envs} of the 256 envs on the host cores"},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
                             "sample": f"{envs} envs per step x {args.steps} steps; oracle/splat_oracle.c (OpenMP, all host threads)"},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 200; 20 for --impl reference)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip other_configs / strong_config3 (profiling runs)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 20 if args.impl == "reference" else 200
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
