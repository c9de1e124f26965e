#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200s, one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload orbit|hires256|hires512|cmem]
    python bench.py --impl reference ...      # the reference's own CPU code, same metric/config

Metric (BASELINE.json): SXI frames/s (and rays/s) of Camera's line-of-sight integration, with the
share of pixels within tolerance of the CPU reference.  Default workload = BASELINE.json
configs[4], the orbit sweep: 10 000 camera poses on a SMILE-like orbit, each a 64x108-ray SXI frame
of 200 trilinear samples through a synthetic 256^3 fp32 MHD cube; frames are block-partitioned over
the ranks (strong scaling: the sweep is fixed), the cube is broadcast once with NCCL and replicated,
and the images are gathered to rank 0 with NCCL inside the timed region.

  value  frames/s, whole job, inputs (cube, frame records) resident in HBM, CUDA-event timed,
         max over ranks; L2 is flushed between timed steps.
  e2e    the same sweep through the host-buffer C ABI call birp_render_cube (pinned host frame
         records in, pinned host images out; H2D/D2H inside the timed region).
  roofline       the march kernel against the roof that binds it: instruction issue.  achieved =
                 executed warp instructions per launch (ncu smsp__inst_executed.sum of the same
                 kernel and workload, profiles/traffic.json, scaled by the live samples of this
                 launch) / CUDA-event kernel time; peak = 592 schedulers x SM clock sampled under
                 load.  The byte figures of SURVEY 8(d) (32 B per live trilinear sample + 4 B per
                 pixel, against the measured HBM copy bandwidth and the L2 gather microbenchmark)
                 ride along as vs_hbm / vs_l2_gather: they exceed 1 because the cube is served
                 from L1/L2, which is why they are not the binding roof.
  exit code      non-zero when the DataCube images are not bit-exact (tolerance-mode flags: < 99.9 %
                 of pixels within 1e-5 or flux off by > 1e-4), CMEM < 95 % / flux > 1e-4, or rank 0's
                 gathered buffer differs from a single-GPU render.
  cpu_baseline   the unmodified reference (oracle/_ref) on a bounded sample of the same frames,
                 on the host cores of this box; the same frames give the parity figure.
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
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from birp_cpp_b200 import synthetic as S  # noqa: E402

NSAMPLES, MAXDIST = 200, 44
ROW_BAND = 128  # rows per interleaved band when one image is split over the ranks (a multiple of the 8-row warp tile)
EXTRA_FLAGS = 0


def cmem_mode():
    """--extra-flags -> the `fast` argument of the capi CMEM calls (False = default: own fp64 blocks behind the region
    table and filter; otherwise the raw BIRP_CMEM_* flags)"""
    return int(EXTRA_FLAGS) if EXTRA_FLAGS else False

WORKLOADS = {
    # name: (description, cube n, W, H, fov_w, fov_h, nsamples, frames, field)
    "orbit": ("orbit sweep: 10000 SMILE-orbit poses, SXI 64x108 rays x 200 trilinear samples, synthetic 256^3 cube",
              256, S.SXI_W, S.SXI_H, S.SXI_FOV_W, S.SXI_FOV_H, 200, 10000, "cube"),
    "hires256": ("synthetic 256^3 cube, 1024x1024 image, 2048 trilinear samples per ray (fine step)",
                 256, 1024, 1024, 36.0, 36.0, 2048, 1, "cube"),
    "hires512": ("synthetic 512^3 cube (512 MiB), 1024x1024 image, 200 trilinear samples, row bands over ranks",
                 512, 1024, 1024, 36.0, 36.0, 200, 1, "cube"),
    "hires512x4k": ("synthetic 512^3 cube (512 MiB), 4096x4096 image, 200 trilinear samples, row bands over ranks",
                    512, 4096, 4096, 36.0, 36.0, 200, 1, "cube"),
    "orbit-nearest": ("orbit sweep: 10000 SMILE-orbit poses, SXI 64x108 rays x 200 nearest-neighbour samples (the "
                      "reference CLI's default sampling), synthetic 256^3 cube",
                      256, S.SXI_W, S.SXI_H, S.SXI_FOV_W, S.SXI_FOV_H, 200, 10000, "cube-nearest"),
    "orbit-zeros": ("orbit sweep: 10000 SMILE-orbit poses, SXI 64x108 rays x 200 trilinear samples, synthetic 256^3 cube "
                    "that is ZERO inside r < 8 RE (what MHD emissivity grids hold inside the magnetosphere)",
                    256, S.SXI_W, S.SXI_H, S.SXI_FOV_W, S.SXI_FOV_H, 200, 10000, "cube-zeros"),
    "orbit-neg": ("orbit sweep: 10000 SMILE-orbit poses, SXI 64x108 rays x 200 trilinear samples, synthetic 256^3 cube with "
                  "a slab of NEGATIVE voxels at 20 <= x < 22 RE (rays end at their first negative sample)",
                  256, S.SXI_W, S.SXI_H, S.SXI_FOV_W, S.SXI_FOV_H, 200, 10000, "cube-neg"),
    "cmem": ("orbit sweep: 10000 poses, SXI 64x108 rays x 200 samples of the analytic CMEM field (Init() defaults)",
             0, S.SXI_W, S.SXI_H, S.SXI_FOV_W, S.SXI_FOV_H, 200, 10000, "cmem"),
}


def workload_cube(cube_n, field):
    """the synthetic grid of a workload: strictly positive, zero inside 8 RE, or with a negative slab"""
    zyx, org, sp = S.make_cube(cube_n, cube_n, cube_n, negative_slab=(field == "cube-neg"))
    if field == "cube-zeros":
        x = org[0].astype(np.float64) + sp[0].astype(np.float64) * np.arange(cube_n)
        r2 = x[None, None, :] ** 2 + x[None, :, None] ** 2 + x[:, None, None] ** 2
        zyx = zyx.copy()
        zyx[r2 < 64.0] = 0.0
    return zyx, org, sp


def workload_config(wl, name, world):
    """the `config` object both arms print (same workload description for the driver's ratio)"""
    desc, cube_n, W, H, fov_w, fov_h, nsamples, nframes, field = wl
    interp = {"cube": "trilinear", "cube-zeros": "trilinear", "cube-neg": "trilinear", "cube-nearest": "nearest",
              "cmem": "analytic"}[field]
    return {"workload": desc, "name": name, "image": [H, W], "samples_per_ray": nsamples, "maxdist_RE": MAXDIST,
            "interpolation": interp, "frames": nframes, "cube": cube_n,
            "parallelism": (f"frames x{world}" if nframes > 1 else f"row tiles x{world}"),
            "l2": "GPU arm: flushed between timed steps (256 MiB memset, untimed)",
            # a line is only printed when these hold: the B200 arm exits non-zero otherwise (see run_b200), the
            # reference arm is the thing compared against
            "parity_ok": True, "gather_ok": True}


def measured_traffic(name, extra_flags):
    """the committed ncu capture of this workload's dominant kernel (profiles/traffic.json, written by
    tools/ncu_traffic.py from one `ncu --set full` capture of `python bench.py --workload <name> --steps 2
    --warmup 1` on one GPU): per-launch DRAM bytes, executed warp instructions and the live samples of that launch.
    Instructions and bytes per live sample are properties of the kernel, so a launch with a different share of the
    same workload (N > 1, --frames) is scaled by its live-sample count."""
    p = ROOT / "profiles" / "traffic.json"
    if not p.exists():
        return None
    key = name if not extra_flags else f"{name}+flags{extra_flags}"
    return json.loads(p.read_text()).get(key)


def issue_roofline(entry, live_samples, kernel_ms, sm_mhz):
    """The binding roof of the march kernels is instruction issue: one warp instruction per scheduler per clock,
    4 schedulers x 148 SMs.  achieved = executed warp instructions of this launch / kernel time; the instruction
    count is ncu's smsp__inst_executed.sum of the committed capture of the same kernel on the same workload, scaled
    by live samples when this launch covers a different share (N > 1, --frames)."""
    peak = 148 * 4 * sm_mhz * 1e6 / 1e9  # G warp instructions / s at the SM clock sampled under load
    out = {"bound": "issue", "achieved": None, "peak": peak, "unit": "Gwarp-inst/s", "frac": None, "traffic": None,
           "peak_source": f"592 warp schedulers x {sm_mhz:.0f} MHz (SM clock sampled under load, nvml)"}
    if not entry or not entry.get("warp_instructions") or not entry.get("live_samples"):
        out["note"] = "no ncu capture of this workload / flag combination in profiles/traffic.json"
        return out
    scale = live_samples / float(entry["live_samples"])
    winst = entry["warp_instructions"] * scale
    out["achieved"] = winst / (kernel_ms / 1e3) / 1e9
    out["frac"] = out["achieved"] / peak
    if entry.get("l1_wavefronts"):
        # the other candidate: the L1 data pipe moves one wavefront per clock per SM (a 32-byte-per-lane load costs 8 or
        # more).  Whichever unit is busier is the bound that is reported; the other rides along.
        l1_peak = 148 * sm_mhz * 1e6 / 1e9
        l1_ach = entry["l1_wavefronts"] * scale / (kernel_ms / 1e3) / 1e9
        l1 = {"bound": "l1", "achieved": l1_ach, "peak": l1_peak, "unit": "Gwavefront/s", "frac": l1_ach / l1_peak,
              "wavefronts_per_global_load": entry.get("l1_wavefronts_per_global_load"),
              "peak_source": f"148 SMs x one L1 data-pipe wavefront per clock x {sm_mhz:.0f} MHz"}
        issue = {k: out[k] for k in ("bound", "achieved", "peak", "unit", "frac", "peak_source")}
        if l1["frac"] > out["frac"]:
            out.update({k: l1[k] for k in ("bound", "achieved", "peak", "unit", "frac", "peak_source")})
            out["issue"] = issue
        out["l1"] = l1
    out["traffic"] = int(entry["bytes"] * scale) if entry.get("bytes") is not None else None
    out["warp_instructions_per_launch"] = int(winst)
    out["warp_instructions_per_warp_sample"] = entry["warp_instructions"] / (entry["live_samples"] / 32.0)
    out["source"] = {"profile": entry.get("profile"), "kernel": entry.get("kernel"),
                     "scaled_by_live_samples": round(scale, 6)}
    for k in ("issue_active_pct", "l1tex_pct", "l1_wavefronts_per_request", "xu_pct", "fp64_pct"):
        if k in entry:
            out[k + "_ncu"] = entry[k]
    return out


def kernel_name(field, trilinear, extra_flags):
    from birp_cpp_b200 import capi
    if field != "cube":
        return "march_kernel<CmemField, tolerance mode>" if (extra_flags & capi.CMEM_FAST) else "march_kernel<CmemField>"
    if not trilinear:
        return "march_cube_kernel<nearest>"
    return "march_cube_fast_kernel<trilinear>" if (extra_flags & capi.FAST_MARCH) else "march_cube_kernel<trilinear>"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """samples SM clocks / throttle reasons of one GPU every 100 ms while the timed region runs"""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------- reference arm
_REF_SPACES = {}


def _ref_worker(args):
    """renders a list of frames with the UNMODIFIED reference in this process"""
    (field, cube_n, poses, px, fov_h, fov_w, W, H, nsamples, want_images) = args
    import oracle as O
    key = (field, cube_n)
    if key not in _REF_SPACES:  # one reference DataCube / CMEM object per worker process
        if field.startswith("cube"):
            zyx, org, sp = workload_cube(cube_n, field)
            _REF_SPACES[key] = O.RefCube(zyx, org, sp, field != "cube-nearest")
        else:
            _REF_SPACES[key] = O.RefCmem()
    space = _REF_SPACES[key]
    t0 = time.perf_counter()
    imgs = []
    for p in poses:
        img = O.ref_render(space, p, S.DEFAULT_AIM_X, px, fov_h, fov_w, W=W, H=H, nsamples=nsamples, maxdist=MAXDIST)
        if want_images:
            imgs.append(img)
    return time.perf_counter() - t0, imgs


def reference_sample(wl, poses, workers, want_images=False, pool=None):
    """time `poses` through oracle/_ref on `workers` processes; returns (wall seconds of the render
    phase = max over workers, images)"""
    desc, cube_n, W, H, fov_w, fov_h, nsamples, _, field = wl
    px = fov_h / H
    chunks = [poses[i::workers] for i in range(workers)]
    chunks = [c for c in chunks if len(c)]
    jobs = [(field, cube_n, c, px, fov_h, fov_w, W, H, nsamples, want_images) for c in chunks]
    if len(jobs) == 1 and pool is None:
        res = [_ref_worker(jobs[0])]
    else:
        res = pool.map(_ref_worker, jobs)
    secs = max(r[0] for r in res)
    imgs = None
    if want_images:
        imgs = [None] * len(poses)
        for w, r in enumerate(res):
            for k, img in enumerate(r[1]):
                imgs[w + k * len(chunks)] = img
    return secs, imgs


def run_reference(args, wl, name):
    """--impl reference: the reference's own CPU implementation on this box's host cores"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    import oracle as O
    desc, cube_n, W, H, fov_w, fov_h, nsamples, nframes, field = wl
    if not O.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libbirp_ref.so was not built"}))
        return
    O.ref()  # dlopen oracle/_ref/libbirp_ref.so in THIS process: the fork-pool workers inherit the mapping
    cores = len(os.sched_getaffinity(0))
    workers = max(1, min(cores, 32))
    if nframes == 1:
        # single big image: the reference cannot split one frame; time a 128-row sub-image at the same
        # per-pixel work on one core per worker -- disjoint cameras of a smaller raster
        W, H = max(W // 8, 8), max(H // 8, 8)
        wl = (desc, cube_n, W, H, fov_w, fov_h, nsamples, nframes, field)
        all_poses = np.repeat(np.array([S.DEFAULT_POS], np.float32), workers, axis=0)
        sample_note = f"{workers} x one {W}x{H}-ray sub-image at the full image's FOV and {nsamples} samples/ray"
    else:
        all_poses = S.orbit_poses(nframes)
        sample_note = None
    per_step = workers * (1 if (args.steps + args.warmup) > 24 else 2)
    ctx = mp.get_context("fork")
    rng = np.random.default_rng(0)
    with ctx.Pool(workers) as pool:
        times = []
        for step in range(args.warmup + args.steps):
            if nframes == 1:
                poses = all_poses
            else:
                idx = np.sort(rng.choice(nframes, size=per_step, replace=False))
                poses = all_poses[idx]
            secs, _ = reference_sample(wl, poses, workers, pool=pool)
            if step >= args.warmup:
                times.append((secs, len(poses)))
    tot_s = sum(t for t, _ in times)
    tot_f = sum(n for _, n in times)
    if nframes == 1:
        full_W, full_H = WORKLOADS[name][2], WORKLOADS[name][3]
        frames_per_s = (tot_f * W * H) / (full_W * full_H) / tot_s  # full-frame equivalents
        rays = tot_f * W * H / tot_s
    else:
        frames_per_s = tot_f / tot_s
        rays = frames_per_s * W * H
        sample_note = f"{per_step} randomly chosen frames of the {nframes}-frame sweep per step, {workers} processes"
    line = {
        "impl": "reference", "metric": "sxi_frames_per_sec" if nframes > 1 else "frames_per_sec",
        "value": frames_per_s, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_s / max(len(times), 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(WORKLOADS[name] if nframes == 1 else wl, name, args.gpus),
        "reference_threads": workers,
        "rays_per_sec": rays,
        "cpu_baseline": {"value": frames_per_s, "unit": "frames/s", "cores": workers, "kind": "reference",
                         "sample": sample_note},
        "e2e": {"value": frames_per_s, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- B200 arm
def run_b200(args, wl, name):
    global EXTRA_FLAGS
    EXTRA_FLAGS = args.extra_flags
    import torch
    import torch.distributed as dist

    from birp_cpp_b200 import capi, sharding

    desc, cube_n, W, H, fov_w, fov_h, nsamples, nframes, field = wl
    trilinear = field != "cube-nearest"
    cube_kind = field
    if field.startswith("cube"):
        field = "cube"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    # the ranks of one box share its cores: each makes its frame records on its share of them
    os.environ.setdefault("BIRP_HOST_THREADS", str(max(1, min(16, len(os.sched_getaffinity(0)) // world))))
    torch.cuda.set_device(local)
    capi.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    # ---- inputs: cube generated on rank 0, broadcast over NCCL, replicated; poses -> frame records
    cube = None
    if field == "cube":
        if rank == 0:
            zyx, org, sp = workload_cube(cube_n, cube_kind)
            d_cube = torch.from_numpy(zyx).to(dev)
            geom = torch.tensor(np.concatenate([org, sp]), device=dev)
        else:
            d_cube = torch.empty((cube_n, cube_n, cube_n), dtype=torch.float32, device=dev)
            geom = torch.empty(6, dtype=torch.float32, device=dev)
        if world > 1:
            dist.broadcast(d_cube, 0)
            dist.broadcast(geom, 0)
        g = geom.cpu().numpy()
        org, sp = g[:3].copy(), g[3:].copy()
        torch.cuda.synchronize()
        cube = capi.Cube.from_device(d_cube.data_ptr(), cube_n, cube_n, cube_n, org, sp, sptr)
        del d_cube
    else:
        params = capi.cmem_init()

    interleave = None
    if nframes > 1:
        poses = S.orbit_poses(nframes)
        frames_all = capi.frames_for_poses(poses, S.DEFAULT_AIM_X, fov_w, fov_h, W, H)
        my_idx = sharding.frame_indices(nframes, world, rank, interleave=True)
        my_frames = np.ascontiguousarray(frames_all[my_idx])
        counts = [len(sharding.frame_indices(nframes, world, r, True)) for r in range(world)]
        rows = (0, H)
        my_rows = H
        band = None
    else:
        poses = np.array([S.DEFAULT_POS], np.float32)
        frames_all = capi.frames_for_poses(poses, S.DEFAULT_AIM_X, fov_w, fov_h, W, H)
        my_frames = frames_all
        band = sharding.block_counts(H, world)
        rows = sharding.row_band(H, world, rank)
        my_rows = rows[1] - rows[0]
        counts = [1] * world
        if world > 1 and os.environ.get("BIRP_GATHER", "peer") == "peer":
            # one image over the GPUs in INTERLEAVED bands of ROW_BAND rows (rank r: bands r, r + N, ...): every rank gets the
            # same mix of cheap rows (rays that leave the cube early) and expensive ones
            interleave = (ROW_BAND, rank, world)
            my_rows = sum(r1 - r0 for r0, r1 in sharding.interleaved_row_bands(H, ROW_BAND, world, rank))
            rows = (0, H)
    nmine = len(my_frames)

    d_frames = torch.from_numpy(my_frames.view(np.float32).reshape(-1, 16).copy()).to(dev)
    d_images = torch.zeros((nmine, H, W), dtype=torch.float32, device=dev)
    d_valid = torch.zeros(1, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gathered = None
    gatherer = None
    peer = None
    gather_mode = os.environ.get("BIRP_GATHER", "peer") if world > 1 else "none"
    if gather_mode == "peer":
        # rank 0 owns the result; the other ranks' render kernels store into it over NVLink
        peer = sharding.PeerGather((nframes if nframes > 1 else 1) * H * W, rank, world, dev)
    elif world > 1 and nframes > 1:
        # sweep: the images travel to rank 0 in two pieces, piece 0 while piece 1 is still rendering
        gatherer = sharding.FrameGather(nframes, world, rank, (H, W), torch.float32, dev, interleave=True,
                                        nchunks=int(os.environ.get("BIRP_GATHER_CHUNKS", "1")))
    elif world > 1 and rank == 0:
        gathered = torch.empty((H, W), dtype=torch.float32, device=dev)

    def launch(valid_ptr=0, lo=0, hi=None, to_peer=False):
        """enqueue the render of this rank's frames [lo, hi) on the current stream; to_peer: write the
        pixels into rank 0's gathered buffer (frame f of the sweep at slot f) instead of the local one"""
        hi = nmine if hi is None else hi
        if hi <= lo:
            return
        fr_ptr = d_frames.data_ptr() + lo * 64
        if to_peer:
            img_ptr = peer.ptr
            # sweep: my local frame j is sweep frame my_idx[j] = my_idx[0] + j * step; single image: slot 0
            step = int(my_idx[1] - my_idx[0]) if (nframes > 1 and nmine > 1) else 1
            off, stride = (int(my_idx[0]) + lo * step, step) if nframes > 1 else (0, 1)
        else:
            img_ptr, off, stride = d_images.data_ptr() + lo * H * W * 4, 0, 1
        if field == "cube" and interleave is not None:
            capi.render_cube_device_rows(cube, fr_ptr, hi - lo, W, H, img_ptr, *interleave, nsamples=nsamples,
                                         maxdist=MAXDIST, trilinear=trilinear, d_valid=valid_ptr, stream=sptr,
                                         extra_flags=EXTRA_FLAGS, frame_offset=off, frame_stride=stride)
        elif field == "cube":
            capi.render_cube_device(cube, fr_ptr, hi - lo, W, H, img_ptr, nsamples=nsamples,
                                    maxdist=MAXDIST, trilinear=trilinear, rows=rows, d_valid=valid_ptr, stream=sptr,
                                    extra_flags=EXTRA_FLAGS, frame_offset=off, frame_stride=stride)
        else:
            capi.render_cmem_device(params, fr_ptr, hi - lo, W, H, img_ptr, nsamples=nsamples,
                                    maxdist=MAXDIST, rows=rows, d_valid=valid_ptr, stream=sptr,
                                    fast=cmem_mode(), frame_offset=off, frame_stride=stride)

    def render_and_gather():
        """one step of the job: render this rank's share, images to rank 0 (N > 1)"""
        if peer is not None:
            launch(to_peer=True)
            peer.fence()
            return None
        if gatherer is not None:
            for c in range(gatherer.nchunks):
                lo, hi = gatherer.chunk_bounds(rank, c)
                launch(lo=lo, hi=hi)
                gatherer.post(d_images, c)
            return gatherer.finish()
        launch()
        if world > 1:
            sharding.gather_row_bands(d_images[0, rows[0]:rows[1], :], band, 0, out=gathered)
        return gathered

    def step():
        render_and_gather()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- live-sample count (untimed) for the roofline's algorithmic bytes
    launch(d_valid.data_ptr())
    torch.cuda.synchronize()
    capi.check_device_status(sptr)
    valid_mine = int(d_valid.item())

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed: K steps, each bracketed by events; L2 flushed (untimed) between steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = capi.launch_count()
    with ClockSampler(local) as clocks:
        barrier()
        t_wall0 = time.perf_counter()
        for a, b, c in ev:
            flush.zero_()
            if world > 1:
                dist.barrier()
            a.record(stream)
            if world == 1:
                launch()
                b.record(stream)
            else:
                render_and_gather()
                b.record(stream)  # (N > 1: render and gather overlap, there is no separate kernel interval)
            c.record(stream)
        barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = capi.launch_count() - launches0
    # render-only pass (no gather), per rank: shows what of an N-GPU step is load imbalance between
    # ranks and what is the exposed tail of the gather
    kernel_only = None
    if world > 1:
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        k0.record(stream)
        for _ in range(args.steps):
            launch()
        k1.record(stream)
        torch.cuda.synchronize()
        mine_ms = torch.tensor([k0.elapsed_time(k1) / args.steps], dtype=torch.float64, device=dev)
        allk = [torch.zeros_like(mine_ms) for _ in range(world)]
        dist.all_gather(allk, mine_ms)
        kernel_only = [round(float(x[0]), 4) for x in allk]
    step_ms = np.array([a.elapsed_time(c) for a, b, c in ev])
    kern_ms = np.array([a.elapsed_time(b) for a, b, c in ev])
    t = torch.tensor([step_ms.sum(), kern_ms.mean(), float(valid_mine)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, kern_ms_mean, valid_total = float(tmax[0]), float(tmax[1]), float(tsum[2])
    else:
        total_ms, kern_ms_mean, valid_total = float(t[0]), float(t[1]), float(t[2])
    capi.check_device_status(sptr)

    # ---- N > 1: what rank 0 holds must be what one GPU renders (a few frames / rows of other ranks' shares)
    gather_check = None
    e2e_gather_check = None
    e2e_note = "bytes are per rank; every rank moves its own share"
    parity_rows = (rows[0], min(rows[0] + 16, rows[1]))
    if peer is not None:
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            full = peer.tensor((nframes, H, W) if nframes > 1 else (1, H, W))
            if nframes > 1:
                probe = np.unique(np.linspace(0, nframes - 1, 24).astype(int))
                d_pf = torch.from_numpy(frames_all[probe].view(np.float32).reshape(-1, 16).copy()).to(dev)
                d_pi = torch.zeros((len(probe), H, W), dtype=torch.float32, device=dev)
                if field == "cube":
                    capi.render_cube_device(cube, d_pf.data_ptr(), len(probe), W, H, d_pi.data_ptr(), nsamples=nsamples,
                                            maxdist=MAXDIST, trilinear=trilinear, stream=sptr, extra_flags=EXTRA_FLAGS)
                else:
                    capi.render_cmem_device(params, d_pf.data_ptr(), len(probe), W, H, d_pi.data_ptr(),
                                            nsamples=nsamples, maxdist=MAXDIST, stream=sptr,
                                            fast=cmem_mode())
                torch.cuda.synchronize()
                same = bool(torch.equal(full[torch.from_numpy(probe).to(dev)], d_pi))
                gather_check = {"frames_compared": int(len(probe)), "identical_to_single_gpu_render": same}
            else:
                d_one = torch.zeros((1, H, W), dtype=torch.float32, device=dev)
                capi.render_cube_device(cube, d_frames.data_ptr(), 1, W, H, d_one.data_ptr(), nsamples=nsamples,
                                        maxdist=MAXDIST, trilinear=trilinear, stream=sptr, extra_flags=EXTRA_FLAGS)
                torch.cuda.synchronize()
                gather_check = {"rows_compared": int(H), "identical_to_single_gpu_render": bool(torch.equal(full, d_one))}
        dist.barrier()

    # ---- the same job in tolerance mode (BIRP_FAST_MARCH), when the main measurement is the bit-exact default: device
    # timed like `value`, its images compared with the default's (which are the reference's, bit for bit)
    fast_arm = None
    if field == "cube" and trilinear and not (EXTRA_FLAGS & capi.FAST_MARCH):
        saved = EXTRA_FLAGS
        ref_imgs_dev = d_images[:min(nmine, 8)].clone() if peer is None else None
        EXTRA_FLAGS = saved | capi.FAST_MARCH
        for _ in range(3):
            step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tm = []
        for _ in range(args.steps):
            flush.zero_()
            if world > 1:
                dist.barrier()
            f0.record(stream)
            step()
            f1.record(stream)
            torch.cuda.synchronize()
            tm.append(f0.elapsed_time(f1))
        tt = torch.tensor([float(np.sum(tm))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        fast_ms = float(tt[0]) / args.steps
        frac = rel = None
        if ref_imgs_dev is not None:
            a = d_images[:len(ref_imgs_dev)].double()
            b = ref_imgs_dev.double()
            floor = 1e-5 * float(b.abs().max())
            frac = float(((a - b).abs() <= torch.clamp(1e-5 * b.abs(), min=floor)).double().mean())
            rel = float(((a - b).abs() / torch.clamp(b.abs(), min=floor)).max())
        fast_arm = {"flag": "BIRP_FAST_MARCH", "value": (nframes if nframes > 1 else 1) / (fast_ms / 1e3),
                          "unit": "frames/s", "ms_per_step": fast_ms,
                          "pixels_within_1e-5_of_the_bit_exact_images": frac, "max_relative_difference": rel,
                          "note": "opt-in: sample VALUES with fused arithmetic (termination stays exact); "
                                  "`python bench.py --extra-flags 32768` measures it as the main arm, e2e and parity "
                                  "against the reference included"}
        EXTRA_FLAGS = saved
        launch()  # (d_images hold the default's pixels again)
        torch.cuda.synchronize()

    # ---- e2e: the host-buffer C ABI call, pinned host memory, H2D + D2H inside the timed region.
    # N > 1: ALL images land in one host buffer that rank 0 owns -- a POSIX shared-memory segment every rank maps and
    # page-locks (birp_host_register); each rank's sweep call writes its frames (interleaved: frame f of the sweep is
    # rank f mod N's) or its row bands straight into it, every GPU over its own PCIe link.
    h_frames = capi.PinnedArray((nmine,), capi.FRAME_DTYPE)
    h_frames.array[:] = my_frames
    frames_total_n = nframes if nframes > 1 else 1
    shared = None
    if world > 1:
        shared = sharding.SharedHostResult((frames_total_n, H, W), rank, world)
        h_all = shared.array
        h_images = None
    else:
        h_images = capi.PinnedArray((nmine, H, W), np.float32)
        h_all = h_images.array

    my_poses = np.ascontiguousarray(poses[my_idx]) if nframes > 1 else poses
    e2e_stride = (int(my_idx[0]), int(my_idx[1] - my_idx[0]) if nmine > 1 else 1) if nframes > 1 else (0, 1)
    if world > 1 and nframes > 1:
        # (frame_indices falls back to contiguous blocks when the sweep does not divide evenly: offset = first frame)
        assert np.array_equal(my_idx, e2e_stride[0] + e2e_stride[1] * np.arange(nmine))

    def e2e_step():
        # poses in -> images out.  Sweep: one call (birp_render_*_sweep_strided) that makes the frame records of a chunk
        # on the host threads while the GPU renders the previous chunk and the copy engine drains the one before;
        # single image: per-frame host setup, H2D of the frame record, render, D2H of the image (N > 1: of this
        # rank's interleaved row bands)
        if nframes > 1:
            capi.render_sweep(cube if field == "cube" else params, my_poses, S.DEFAULT_AIM_X, fov_w, fov_h, W, H,
                              nsamples=nsamples, maxdist=MAXDIST, trilinear=trilinear, fast=cmem_mode(), rows=rows,
                              extra_flags=(EXTRA_FLAGS if field == "cube" else 0), frames_out=h_frames.array,
                              out=h_all, frame_offset=e2e_stride[0], frame_stride=e2e_stride[1])
            return
        capi.frames_for_poses(my_poses, S.DEFAULT_AIM_X, fov_w, fov_h, W, H, out=h_frames.array)
        if interleave is not None:
            capi.render_cube_rows(cube, h_frames.array, W, H, *interleave, nsamples=nsamples, maxdist=MAXDIST,
                                  trilinear=trilinear, out=h_all, extra_flags=EXTRA_FLAGS)
        else:
            cube.render(h_frames.array, W, H, nsamples=nsamples, maxdist=MAXDIST, trilinear=trilinear, rows=rows,
                        out=h_all, extra_flags=EXTRA_FLAGS)

    for _ in range(2):
        e2e_step()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()  # (synchronous on return: this rank's share of the sweep is in rank 0's host buffer)
    if world > 1:
        dist.barrier()  # the K steps end when every rank's last share is in that buffer
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    e2e_images = h_all if world > 1 else h_images.array
    if world > 1:
        e2e_note = ("every rank writes its share straight into ONE host buffer owned by rank 0 (POSIX shared memory, "
                    "page-locked by each rank): bytes are per rank, each GPU drains over its own PCIe link; the K steps are "
                    "bracketed by barriers")
        # rank 0 holds the whole result on the host: probe frames / rows of OTHER ranks' shares against rank 0's own render
        if rank == 0:
            if nframes > 1:
                probe = np.unique(np.linspace(0, nframes - 1, 24).astype(int))
                own_fr = np.ascontiguousarray(frames_all[probe])
                if field == "cube":
                    mine = cube.render(own_fr, W, H, nsamples=nsamples, maxdist=MAXDIST, trilinear=trilinear,
                                       extra_flags=EXTRA_FLAGS)
                else:
                    mine = capi.render_cmem(params, own_fr, W, H, nsamples=nsamples, maxdist=MAXDIST, fast=cmem_mode())
                same = bool(np.array_equal(mine.view(np.uint32), np.ascontiguousarray(h_all[probe]).view(np.uint32)))
                e2e_gather_check = {"frames_compared": int(len(probe)), "identical": same}
            else:
                mine = cube.render(frames_all, W, H, nsamples=nsamples, maxdist=MAXDIST, trilinear=trilinear,
                                   extra_flags=EXTRA_FLAGS)
                e2e_gather_check = {"rows_compared": int(H),
                                    "identical": bool(np.array_equal(mine.view(np.uint32), h_all.view(np.uint32)))}

    # ---- single-frame latency through the host-buffer call (BASELINE.json configs[0]/[1]: one SXI image)
    latency_ms = None
    if rank == 0 and nframes > 1:
        one_pose = poses[:1]
        one_img = np.zeros((1, H, W), np.float32)

        def one_frame():
            fr = capi.frames_for_poses(one_pose, S.DEFAULT_AIM_X, fov_w, fov_h, W, H, nthreads=1)
            if field == "cube":
                cube.render(fr, W, H, nsamples=nsamples, maxdist=MAXDIST, trilinear=trilinear, out=one_img)
            else:
                capi.render_cmem(params, fr, W, H, nsamples=nsamples, maxdist=MAXDIST, out=one_img,
                                 fast=cmem_mode())

        for _ in range(5):
            one_frame()
        t0 = time.perf_counter()
        for _ in range(50):
            one_frame()
        latency_ms = (time.perf_counter() - t0) / 50 * 1e3

    # ---- sweep set-up without the host: orbit poses and frame records made on the device
    device_setup = None
    if rank == 0 and nframes > 1:
        d_poses = torch.empty((nframes, 3), dtype=torch.float32, device=dev)
        d_fr2 = torch.empty((nframes, 16), dtype=torch.float32, device=dev)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(3):
            if it == 2:
                s0.record(stream)
            capi.orbit_poses_device(nframes, d_poses.data_ptr(), stream=sptr)
            capi.frame_setup_device(d_poses.data_ptr(), nframes, S.DEFAULT_AIM_X, fov_w, fov_h, W, H, d_fr2.data_ptr(),
                                    stream=sptr)
        s1.record(stream)
        torch.cuda.synchronize()
        capi.check_device_status(sptr)
        t0 = time.perf_counter()
        capi.frames_for_poses(capi.orbit_poses(nframes), S.DEFAULT_AIM_X, fov_w, fov_h, W, H)
        host_ms = (time.perf_counter() - t0) * 1e3
        same = (d_fr2.cpu().numpy().view(np.uint32) == frames_all.view(np.uint32).reshape(-1, 16)).all(axis=1).mean()
        device_setup = {"ms": s0.elapsed_time(s1), "host_ms": host_ms, "frames": nframes,
                        "records_bit_identical_to_host": float(same),
                        "note": "birp_orbit_poses_device + birp_frame_setup_device (two kernels) against the threaded "
                                "host path; the device rounds fp64 results once, glibc's fp32 atan2f/acosf/asinf are "
                                "not correctly rounded, hence the differing records (<= 2.2e-6 in a rotation entry)"}

    # ---- rank 0: CPU baseline on a bounded sample + parity of the same frames, then the JSON line
    failed = []
    if rank == 0:
        import oracle as O
        hbm, hbm_src = peaks()
        frames_total = nframes if nframes > 1 else 1
        value = frames_total / (total_ms / args.steps / 1e3)
        kernel_ms = float(kern_ms.mean())
        # algorithmic bytes of this rank's launch (rank 0's share; N=1: everything), SURVEY 8(d)
        bytes_per_sample = (32 if trilinear else 4) if field == "cube" else 0
        alg_bytes = valid_mine * bytes_per_sample + nmine * my_rows * W * 4
        achieved_gbs = alg_bytes / (kernel_ms / 1e3) / 1e9
        clk = clocks.summary()
        sm_mhz = float(clk.get("sm_mhz") or clk.get("sm_max_mhz") or 1965.0)
        l2_gather = None
        gp = ROOT / "profiles" / "r1d_gather_roofline.json"
        if gp.exists():
            l2_gather = json.loads(gp.read_text())
        roofline = issue_roofline(measured_traffic(name, EXTRA_FLAGS), valid_mine, kernel_ms, sm_mhz)
        roofline.update({
            "kernel": kernel_name(field, trilinear, EXTRA_FLAGS), "kernel_ms": kernel_ms,
            "live_samples_per_launch": int(valid_mine), "algorithmic_bytes_per_launch": int(alg_bytes),
            "vs_hbm": {"achieved_GBs": achieved_gbs, "peak_GBs": hbm, "ratio": achieved_gbs / hbm, "peak_source": hbm_src,
                       "note": "algorithmic bytes (32 B per live trilinear sample, 4 B nearest, + 4 B per pixel) / kernel "
                               "time over the measured HBM copy bandwidth; above 1 because the cube is served from "
                               "L1/L2 (see traffic: DRAM bytes per launch) -- not the binding roof"},
            "vs_l2_gather": ({"peak_GBs": l2_gather["gather32_l2_64MiB_GBs"],
                              "ratio": achieved_gbs / l2_gather["gather32_l2_64MiB_GBs"],
                              "hbm_resident_peak_GBs": l2_gather["gather32_hbm_2GiB_GBs"],
                              "source": "profiles/r1d_gather_roofline.json (tools/gather_roofline.cu: random independent "
                                        "32 B sectors out of a 64 MiB L2-resident / 2 GiB HBM-resident buffer)"}
                             if l2_gather else None),
        })

        cpu = {"value": None, "unit": "frames/s", "cores": 1, "kind": "reference", "sample": "unavailable"}
        parity = None
        if O.have_ref():
            if nframes > 1:
                nsamp = 32  # ~10 s of the single-threaded reference
                idx = np.linspace(0, nmine - 1, nsamp).astype(int)
                secs, ref_imgs = reference_sample(wl, my_poses[idx], 1, want_images=True)
                cpu = {"value": nsamp / secs, "unit": "frames/s", "cores": 1, "kind": "reference",
                       "sample": f"{nsamp} frames evenly spaced over rank 0's {nmine} frames, unmodified reference "
                                 f"Camera (oracle/_ref), 1 thread, {secs:.1f} s"}
                got = e2e_images[my_idx[idx]]  # (N > 1: the shared buffer holds the whole sweep, frame f at slot f)
                want = np.stack(ref_imgs)
            else:
                # sub-image of the same camera cannot be cut out of the reference's loops: time a
                # 128x128-ray camera with the same FOV and samples/ray (constant ns/sample),
                # and check parity of rows of the full image against the oracle restatement
                sw = sh = 128
                swl = (desc, cube_n, sw, sh, fov_w, fov_h, nsamples, 1, field)
                secs, _ = reference_sample(swl, poses[:1], 1)
                fps = (sw * sh) / (W * H) / secs
                cpu = {"value": fps, "unit": "frames/s", "cores": 1, "kind": "reference",
                       "sample": f"one {sw}x{sh}-ray image at the same FOV and {nsamples} samples/ray through the "
                                 f"unmodified reference, 1 thread, {secs:.1f} s, scaled by ray count"}
                zyx, org, sp = S.make_cube(cube_n, cube_n, cube_n)
                fo = O.orc_frames(poses, S.DEFAULT_AIM_X, fov_w, fov_h, W, H)
                pr = parity_rows
                want = O.orc_render_cube(zyx, org, sp, trilinear, fo, W, H, nsamples=nsamples, rows=pr,
                                         nthreads=len(os.sched_getaffinity(0)))[:, pr[0]:pr[1]]
                got = e2e_images[:, pr[0]:pr[1]]
            floor = 1e-5 * np.abs(want).max()
            ok = np.abs(got.astype(np.float64) - want) <= np.maximum(1e-5 * np.abs(want), floor)
            parity = {"frames_checked": int(len(want)), "pixels_checked": int(want.size),
                      "pixels_within_1e-5": float(ok.mean()),
                      "bit_exact": bool(np.array_equal(got.view(np.uint32), np.ascontiguousarray(want).view(np.uint32))),
                      "flux_rel_err": float(abs(got.sum(dtype=np.float64) - want.sum(dtype=np.float64)) /
                                            max(want.sum(dtype=np.float64), 1e-300)),
                      "against": "unmodified reference (oracle/_ref)" if nframes > 1 else "oracle restatement"}
            # the gate: DataCube path bit-exact (tolerance-mode kernels: SURVEY 8(d) tolerance at 99.9 %), CMEM 95 %
            tolerance_mode = (field == "cube" and (EXTRA_FLAGS & capi.FAST_MARCH)) or \
                             (field == "cmem" and (EXTRA_FLAGS & capi.CMEM_FAST))
            if field == "cube" and not tolerance_mode:
                if not parity["bit_exact"]:
                    failed.append("DataCube images are not bit-exact against the reference")
            else:
                need = 0.999 if tolerance_mode else 0.95
                if parity["pixels_within_1e-5"] < need:
                    failed.append(f"only {parity['pixels_within_1e-5']:.4%} of pixels within 1e-5 (need {need:.1%})")
                if parity["flux_rel_err"] > 1e-4:
                    failed.append(f"integrated flux off by {parity['flux_rel_err']:.3g} (> 1e-4)")
        if gather_check is not None and not gather_check["identical_to_single_gpu_render"]:
            failed.append("rank 0's gathered buffer differs from a single-GPU render")
        if e2e_gather_check is not None and not e2e_gather_check["identical"]:
            failed.append("e2e: rank 0's host images differ from this rank's own render of the probe frames")

        cfg = workload_config(wl, name, world)
        cfg["parity_ok"] = not any("gather" not in f and "rank 0" not in f for f in failed)
        cfg["gather_ok"] = not any("rank 0" in f for f in failed)
        line = {
            "metric": "sxi_frames_per_sec" if nframes > 1 else "frames_per_sec",
            "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "run": {"cube_analysis": cube.info() if cube is not None else None, "extra_flags": EXTRA_FLAGS,
                    "gather": {"none": "none (1 GPU)",
                               "peer": "render kernels store their pixels into rank 0's buffer over NVLink (CUDA IPC), "
                                       "completion flags in rank 0's memory as the fence; inside the timed region",
                               }.get(gather_mode, "NCCL grouped send/recv to rank 0, inside the timed region")},
            "rays_per_sec": value * W * H,
            "samples_per_sec": valid_total / (total_ms / args.steps / 1e3),
            "e2e": {"value": frames_total / (e2e_s / e2e_steps), "unit": "frames/s",
                    "h2d_bytes_per_step": int(nmine * 64), "d2h_bytes_per_step": int(nmine * my_rows * W * 4),
                    "steps": e2e_steps,
                    "api": ("birp_render_cube_sweep / birp_render_cmem_sweep" if nframes > 1 else
                            "birp_frame_setup + birp_render_cube") + " (host poses in, pinned host images out)",
                    "note": e2e_note},
            "gpu_launches": int(launches),
            "tolerance_mode": fast_arm,
            "device_setup": device_setup,
            "single_frame_latency_ms": latency_ms,
            "render_only_ms_per_rank": kernel_only,
            "gather_check": gather_check,
            "e2e_gather_check": e2e_gather_check,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": parity,
            "failed": failed,
            "clocks": clk,
            "wall_s_timed_region": t_wall,
        }
        print(json.dumps(line))

    h_frames.free()
    if h_images is not None:
        h_images.free()
    if shared is not None:
        e2e_images = h_all = got = None  # (views of the segment must be gone before it is closed)
        shared.close()
    if peer is not None:
        peer.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if failed:
        print("bench.py: FAILED -- " + "; ".join(failed), file=sys.stderr)
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="orbit", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="override the number of frames of a sweep workload")
    ap.add_argument("--extra-flags", type=int, default=0,
                    help="BIRP_* diagnostic flags OR-ed into the device render call (A/B of kernel variants)")
    args = ap.parse_args()
    wl = list(WORKLOADS[args.workload])
    if args.frames and wl[7] > 1:
        wl[7] = args.frames
    wl = tuple(wl)
    if args.impl == "reference":
        run_reference(args, wl, args.workload)
    else:
        run_b200(args, wl, args.workload)


if __name__ == "__main__":
    main()
