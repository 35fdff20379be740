#!/usr/bin/env python
"""bench.py -- Mrays/s (primary+shadow+disc) and ms/frame of the octree hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): synthetic ~90 M-point "abandoned-scale" static level +
~10 M-point figure in the dynamic tree, depth-12 octree over the 1800-unit cube, one
1920x1080 primary+shadow frame per step, four fixed camera poses cycled step by step.
A step = octree_glc_update() of one frame.  N > 1 (torchrun, one rank per GPU): the octree
is replicated, the image is split into interleaved 64x64 tiles, every rank renders its
tiles straight into rank 0's framebuffer over NVLink (peer stores from the render kernel)
and one small NCCL all-reduce fences the frame.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
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

WIDTH, HEIGHT = 1920, 1080
MAXLEVEL, BASESIZE = 12, 1800.0
TILE = 64
DIVISION = 0
METRIC = "Mrays/s (primary+shadow) at 1080p, 90M-pt octree"
UNIT = "Mrays/s"


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------
# scene: generated once per box, cached as raw .npy so that the N=1,2,4,8 runs
# (and the ranks of one run) share it
# ---------------------------------------------------------------------------
_FIELDS = ("pnt_s", "col_s", "nrm_s", "oct_s", "pnt_d", "col_d", "nrm_d", "oct_d")


def _cache_dir(scale):
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    return os.path.join(base, "qb_scene_c2_v3_%s" % ("%.4f" % scale).replace(".", "p"))


def get_scene(scale, rank, barrier):
    from qubatron_b200 import scene as S
    d = _cache_dir(scale)
    done = os.path.join(d, "DONE.json")
    if rank == 0 and not os.path.exists(done):
        t = time.time()
        sc = S.make_c2(scale=scale, progress=None)
        os.makedirs(d, exist_ok=True)
        for f in _FIELDS:
            np.save(os.path.join(d, f + ".npy"), getattr(sc, f))
        meta = sc.describe()
        meta["cameras"] = sc.cameras
        with open(done + ".tmp", "w") as fh:
            json.dump(meta, fh)
        os.replace(done + ".tmp", done)
        log("scene generated in %.1f s -> %s" % (time.time() - t, d))
    barrier()
    meta = json.load(open(done))
    arrs = {f: np.load(os.path.join(d, f + ".npy"), mmap_mode="r") for f in _FIELDS}
    sc = S.Scene(meta["name"], arrs["pnt_s"], arrs["col_s"], arrs["nrm_s"], arrs["oct_s"], arrs["pnt_d"],
                 arrs["col_d"], arrs["nrm_d"], arrs["oct_d"], meta["basesize"], meta["levels"],
                 meta["static_points_raw"], meta["dynamic_points_raw"])
    sc.cameras = [(tuple(p), tuple(a)) for p, a in meta["cameras"]]
    return sc, meta


# ---------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the render kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------
# CPU legs
# ---------------------------------------------------------------------------
def cpu_port_sample(sc, pose, rows, threads):
    """oracle restatement (full pixel program, both trees), `rows` rows of the 1080p frame"""
    from oracle import qb_oracle as O
    osc = O.OracleScene(sc)
    u = O.uniforms(WIDTH, HEIGHT, pose[0], pose[1], maxlevel=MAXLEVEL, basesize=BASESIZE)
    t = time.time()
    r = O.render(osc, u, rows=rows, threads=threads, want_aux=False, div=DIVISION)
    dt = time.time() - t
    c = r["counters"]
    rays = c["rays_primary"] + c["rays_shadow"] + c["rays_disc"]
    return rays, dt


def pick_rows(nrows):
    """rows spread over the frame (every stride-th band of 4) so the sample sees sky and geometry alike"""
    bands = max(1, nrows // 4)
    stride = HEIGHT // bands
    return [(b * stride, b * stride + 4) for b in range(bands)]


def cpu_port_bands(sc, pose, nrows, threads):
    rays, dt = 0, 0.0
    for r0, r1 in pick_rows(nrows):
        a, b = cpu_port_sample(sc, pose, (r0, min(r1, HEIGHT)), threads)
        rays += a
        dt += b
    return rays, dt


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.

    oracle/_ref/libqubatron_ref.so is the reference's octree.c compiled unmodified; its octree_trace_line
    (octree.c L341-537) is the CPU twin of the shader's traversal and what the engine itself uses on the CPU
    (qubatron.c L268-269 traces the static and the dynamic tree one after the other).  A step traces a bounded
    sample of the frame's rays: for sampled pixels the primary ray in both trees, then the shadow ray from the
    light to the nearer hit in both trees.  Falls back to the oracle port when oracle/_ref is absent."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import qb_oracle as O
    sc, meta = get_scene(args.scale, 0, lambda: None)
    cores = os.cpu_count() or 1
    poses = sc.cameras
    use_ref = O.have_ref()
    step_rays, step_s = [], []
    if use_ref:
        ts = O.RefOctree(BASESIZE, MAXLEVEL)
        ts.insert_points(np.asarray(sc.pnt_s))
        td = O.RefOctree(BASESIZE, MAXLEVEL)
        td.insert_points(np.asarray(sc.pnt_d))
        assert len(ts) == len(sc.oct_s) and len(td) == len(sc.oct_d)
        rng = np.random.default_rng(99)
        nsample = args.ref_pixels

        def one_step(i):
            pos, ang = poses[i % len(poses)]
            u = O.uniforms(WIDTH, HEIGHT, pos, ang, maxlevel=MAXLEVEL, basesize=BASESIZE)
            # sampled 8x4 pixel patches (the GPU's warp footprint), whole frame covered in expectation
            n_patch = nsample // 32
            bx = rng.integers(0, WIDTH // 8, n_patch) * 8
            by = rng.integers(0, HEIGHT // 4, n_patch) * 4
            xs = (bx[:, None] + (np.arange(32) % 8)[None, :]).ravel()
            ys = (by[:, None] + (np.arange(32) // 8)[None, :]).ravel()
            import ctypes as C
            d = (C.c_float * 3)()
            dirs = np.empty((len(xs), 3), np.float32)
            for k in range(len(xs)):
                O.lib().qb_oracle_pixel_ray(C.byref(u), int(xs[k]), int(ys[k]), C.cast(d, C.c_void_p))
                dirs[k] = (d[0], d[1], d[2])
            org = np.tile(np.asarray(pos, np.float32), (len(xs), 1))
            light = np.array([u.light[0], u.light[1], u.light[2]], np.float32)
            t = time.time()
            i_s, tlf_s = ts.trace(org, dirs, threads=0)
            i_d, tlf_d = td.trace(org, dirs, threads=0)
            hit = (i_s != 0) | (i_d != 0)
            # shadow ray towards the centre of the hit leaf (the CPU twin returns the leaf cube, not the isp)
            tl = np.where((i_s != 0)[:, None], tlf_s, tlf_d)[hit]
            tgt = np.stack([tl[:, 0] + tl[:, 3] * 0.5, tl[:, 1] - tl[:, 3] * 0.5, tl[:, 2] - tl[:, 3] * 0.5], axis=1)
            sdir = (tgt - light[None, :]).astype(np.float32)
            sorg = np.tile(light, (len(sdir), 1))
            if len(sdir):
                ts.trace(sorg, sdir, threads=0)
                td.trace(sorg, sdir, threads=0)
            dt = time.time() - t
            return len(xs) + len(sdir), dt
        kind, sample = "reference", ("octree_trace_line (oracle/_ref, octree.c unmodified), %d sampled pixels/step "
                                     "(8x4 patches): primary + shadow ray, static and dynamic tree each, "
                                     "OpenMP over rays" % nsample)
    else:
        def one_step(i):
            return cpu_port_bands(sc, poses[i % len(poses)], args.cpu_rows, 0)
        kind, sample = "port", "oracle restatement, %d rows of the 1080p frame per step, OpenMP over rows" % args.cpu_rows
    for i in range(args.warmup):
        one_step(i)
    for i in range(args.steps):
        r, s = one_step(i)
        step_rays.append(r)
        step_s.append(s)
    total_s = float(np.sum(step_s))
    value = float(np.sum(step_rays)) / total_s / 1e6
    extras = {}
    if not args.no_c1:
        extras["c1_640x360_reference_shader_on_llvmpipe"] = c1_llvmpipe()
    out = {"impl": "reference", "extras": extras, "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": bench_config(meta, args, 1),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(json.dumps(out))


def c1_llvmpipe():
    """BASELINE configs[0]: the reference's UNMODIFIED shader on Mesa llvmpipe (oracle/_ref/glsl_ref), 1 M-point
    cloud, one 640x360 primary+shadow frame; steady-state frame time with all host threads and with one."""
    from oracle import qb_oracle as O
    from qubatron_b200 import scene as S
    if not O.have_glsl():
        return {"unavailable": "oracle/_ref/glsl_ref or the bundled Mesa libGL is missing on this box"}
    try:
        sc = S.make_c1()
        u = O.uniforms(640, 360, *S.CAMERA_C1)
        c = O.render(O.OracleScene(sc), u, want_aux=False)["counters"]
        rays = c["rays_primary"] + c["rays_shadow"] + c["rays_disc"]
        _, info = O.glsl_render(sc, u, repeat=6)
        steady = float(np.median(info["frame_s"][1:]))
        _, info1 = O.glsl_render(sc, u, repeat=2, threads=1)
        return {"renderer": info["renderer"], "rays_per_frame": rays, "first_frame_s": info["frame_s"][0],
                "frame_s": steady, "mrays_s": rays / steady / 1e6, "threads": "LP_NUM_THREADS default (= cores, max 16)",
                "frame_s_1_thread": info1["frame_s"][-1], "mrays_s_1_thread": rays / info1["frame_s"][-1] / 1e6,
                "cores": os.cpu_count()}
    except Exception as e:  # the baseline is reported, never fatal
        return {"unavailable": repr(e)[:300]}


def c1_gpu(K, device):
    """The same configs[0] frame on the GPU through the C ABI (device time per frame)."""
    from oracle import qb_oracle as O
    from qubatron_b200 import scene as S
    sc = S.make_c1()
    rc = K.OctreeGlc(b"", device=device)
    rc.upload_scene(sc)
    rc.enable_counters(True)
    rc.update(640, 360, *S.CAMERA_C1)
    c = rc.read_counters()
    rc.enable_counters(False)
    ms = []
    for _ in range(24):
        rc.update(640, 360, *S.CAMERA_C1)
        ms.append(rc.last_frame_ms())
    rc.destroy()
    rays = c["rays_primary"] + c["rays_shadow"] + c["rays_disc"]
    m = float(np.median(ms[4:]))
    return {"rays_per_frame": rays, "frame_ms": m, "mrays_s": rays / m / 1e3,
            "note": "1 M-point cloud, 640x360, octree (28 MB) is L2-resident; kernel time, CUDA events"}


def c5_views(rc, sc, rank, world, dev, dist, torch):
    """BASELINE configs[4]: 64 views with random (incoherent) cameras on the same level, whole views sharded
    across the ranks (no exchange at all), one octree_cuc_update_views launch per rank."""
    rng = np.random.default_rng(777)
    n = 64
    pos = np.stack([rng.uniform(150, 1650, n), rng.uniform(90, 330, n), rng.uniform(150, 1650, n)], axis=1)
    ang = np.stack([rng.uniform(0, 2 * np.pi, n), rng.uniform(-0.6, 0.6, n), np.zeros(n)], axis=1)
    mine = list(range(rank, n, world))
    rc.set_frame_target(0, 0)
    rc.set_shard(0, 1, TILE, TILE)
    rc.enable_counters(True)
    rc.update_views(WIDTH, HEIGHT, pos[mine], ang[mine], 0.0, 10, MAXLEVEL, BASESIZE, 0)
    c = rc.read_counters()
    rc.enable_counters(False)
    rays = torch.tensor([c["rays_primary"] + c["rays_shadow"] + c["rays_disc"]], dtype=torch.int64, device=dev)
    ms = []
    for _ in range(3):
        rc.update_views(WIDTH, HEIGHT, pos[mine], ang[mine], 0.0, 10, MAXLEVEL, BASESIZE, 0)
        ms.append(rc.last_frame_ms())
    t = torch.tensor([float(np.median(ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(rays)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"views": n, "views_per_gpu": len(mine), "rays": int(rays.item()), "batch_ms": float(t.item()),
            "ms_per_view": float(t.item()) / n, "mrays_s": int(rays.item()) / float(t.item()) / 1e3,
            "note": "device time of the slowest rank for its share of the 64 views (1080p each)"}


def bench_config(meta, args, world):
    return {"workload": "C2: %s -- %d static points (%d octree nodes) + %d dynamic points (%d nodes), depth %d, "
                        "cube %.0f, %dx%d primary+shadow(+light-disc) frame, %d camera poses cycled"
                        % (meta["name"], meta["static_points"], meta["static_nodes"], meta["dynamic_points"],
                           meta["dynamic_nodes"], MAXLEVEL, BASESIZE, WIDTH, HEIGHT, len(meta["cameras"])),
            "scale": args.scale, "width": WIDTH, "height": HEIGHT, "maxlevel": MAXLEVEL,
            "parallelism": "image tiles %dx%d interleaved over %d GPU(s), octree replicated" % (args.tile, args.tile, world),
            "tile_order": "heaviest first, from the tile costs measured the last time the same view was rendered "
                          "(each of the cycled poses is first seen in warm-up); the ordering kernel runs inside the "
                          "timed step; QB_TILE_FEEDBACK=0 renders tiles in image order",
            "l2": "flushed before every step (256 MiB memset, outside the step's event pair); scene arrays "
                  "(%.1f GB) also exceed L2" % ((meta["static_nodes"] + meta["dynamic_nodes"]) * 36e-9
                                                + (meta["static_points"] + meta["dynamic_points"]) * 32e-9)}


# ---------------------------------------------------------------------------
# main arm
# ---------------------------------------------------------------------------
class _CleanStdout:
    """Libraries (NCCL's version banner, for one) write to stdout; the contract is ONE JSON line there.
    Everything printed while this is active goes to stderr, the JSON line is written to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)


OUT = None


def emit(line):
    if OUT is not None:
        OUT.emit(line)
    else:
        print(line, flush=True)


def main():
    global OUT
    with _CleanStdout() as OUT:
        _main()
    OUT = None


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--scale", type=float, default=1.0, help="1.0 = the full 90M-point level")
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 generic, 2 fast")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--division", type=int, default=0, help="0 GLSL a*(1/b) (reference shader on llvmpipe), 1 IEEE")
    ap.add_argument("--cpu-rows", type=int, default=24, help="rows of the frame the cpu_baseline sample renders")
    ap.add_argument("--ref-pixels", type=int, default=65536)
    ap.add_argument("--cpu-passes", type=int, default=10)
    ap.add_argument("--tile", type=int, default=64, help="shard tile edge in pixels (multiple of 16)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-c1", action="store_true", help="skip the configs[0] (C1, 640x360) side measurement")
    ap.add_argument("--c5", action="store_true", help="also measure configs[4]: 64 random views sharded by view")
    ap.add_argument("--res", default="1080p", choices=["1080p", "4k"],
                    help="4k = BASELINE configs[2] (3840x2160, tile split); the default is the metric's 1080p")
    args = ap.parse_args()

    global DIVISION, WIDTH, HEIGHT, METRIC
    DIVISION = args.division
    if args.res == "4k":
        WIDTH, HEIGHT = 3840, 2160
        METRIC = "Mrays/s (primary+shadow) at 2160p, 90M-pt octree"
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    from qubatron_b200 import connector as K
    sc, meta = get_scene(args.scale, rank, barrier)
    poses = sc.cameras

    # a real (non-legacy) stream shared by the connector, torch's memsets and NCCL
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    rc = K.OctreeGlc(b"", device=local)
    rc.set_stream(stream.cuda_stream)
    t0 = time.time()
    rc.upload_scene(sc)
    rc.sync()
    log("rank %d: scene uploaded in %.1f s, device memory %.2f GB" % (rank, time.time() - t0, rc.memsize / 1e9))
    rc.set_kernel(args.kernel)
    rc.set_division(args.division)

    def render(i):
        pos, ang = poses[i % len(poses)]
        rc.update(WIDTH, HEIGHT, pos, ang, 0.0, 10, MAXLEVEL, BASESIZE, 0)

    # frame assembly for N > 1 (qubatron_b200/multigpu.py): peer stores into rank 0's framebuffer, or NCCL reduce
    from qubatron_b200 import multigpu
    sharded = multigpu.ShardedFrame(rc, WIDTH, HEIGHT, rank, world, dev, gather=args.gather, tile=args.tile)
    assemble = sharded.assemble

    # ---- per-pose work counters (counting instantiation, outside the timed region) ----
    rc.enable_counters(True)
    per_pose = []
    for i in range(len(poses)):
        render(i)
        c = rc.read_counters()
        t = torch.tensor([c[k] for k in sorted(c)], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        per_pose.append(dict(zip(sorted(c), [int(v) for v in t.tolist()])))
    rc.enable_counters(False)
    kernel_used = rc.last_kernel()

    def rays_of(c):
        return c["rays_primary"] + c["rays_shadow"] + c["rays_disc"]

    def alg_bytes(c):
        return 32 * (c["expand_s"] + c["expand_d"]) + 4 * (c["leaf_s"] + c["leaf_d"]) + 24 * c["hits"] + 4 * WIDTH * HEIGHT

    flush_buf = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush():
        if flush_buf is not None:
            flush_buf.fill_(1)

    # ---- warm-up -------------------------------------------------------------------
    for i in range(args.warmup):
        flush()
        render(i)
        assemble()
    torch.cuda.synchronize()
    barrier()

    # ---- timed region: K steps, device-timed per step (events on the launch stream) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = rc.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = []
    torch.cuda.synchronize()
    barrier()
    wall0 = time.time()
    for i in range(args.steps):
        flush()
        ev[i][0].record(stream)
        render(i)
        assemble()
        ev[i][1].record(stream)
        kern_ms.append(None)
    torch.cuda.synchronize()
    barrier()
    wall = time.time() - wall0
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)  # max over ranks, step by step
    step_ms = step_ms.cpu().numpy()
    launches = rc.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # kernel-only duration for the roofline: the connector's own event pair brackets just the render kernel
    kms = []
    for i in range(min(args.steps, 12)):
        flush()
        render(i)
        kms.append((i % len(poses), rc.last_frame_ms()))
    kt = torch.tensor([m for _, m in kms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    kt = kt.cpu().numpy()

    # ---- e2e: the reference-facing call with host buffers: pose in (pinned constants H2D inside
    # octree_glc_update), frame out (D2H into host memory) every step
    e2e = None
    if world == 1:
        # two page-locked host frames: the copy of frame i (octree_cuc_read_frame_async, own copy stream, second
        # device framebuffer) overlaps the rendering of frame i+1; every step still moves its frame to the host
        host_frames = [np.empty((HEIGHT, WIDTH, 4), dtype=np.uint8) for _ in range(2)]
        for hf in host_frames:
            torch.cuda.cudart().cudaHostRegister(hf.ctypes.data, hf.nbytes, 0)
        host_frame = host_frames[0]
        for i in range(3):
            render(i)
            rc.read_frame_async(host_frames[i & 1])
        rc.wait_reads()
        torch.cuda.synchronize()
        t = time.time()
        for i in range(args.steps):
            flush()
            render(i)
            rc.read_frame_async(host_frames[i & 1])
        rc.wait_reads()
        torch.cuda.synchronize()
        e2e_s = time.time() - t
        e2e_rays = sum(rays_of(per_pose[i % len(poses)]) for i in range(args.steps))
        # the flush is not part of the call: subtract its measured cost
        tf = 0.0
        if flush_buf is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(args.steps):
                flush()
            b.record(stream)
            torch.cuda.synchronize()
            tf = a.elapsed_time(b) / 1e3
        e2e = {"value": e2e_rays / max(e2e_s - tf, 1e-9) / 1e6, "unit": UNIT, "h2d_bytes_per_step": 88,
               "d2h_bytes_per_step": int(host_frame.nbytes), "ms_per_step": 1e3 * (e2e_s - tf) / args.steps}
        e2e["note"] = "frame i's device-to-host copy overlaps the rendering of frame i+1 (read_frame_async)"
        for hf in host_frames:
            torch.cuda.cudart().cudaHostUnregister(hf.ctypes.data)
    else:
        # rank 0 reads the assembled frame back EVERY step into one of two page-locked host frames; the copy of
        # frame i (device-to-device into a staging buffer after the fence, then to the host on the copy stream)
        # overlaps the rendering of frame i+1
        if rank == 0:
            host_frames = [np.empty((HEIGHT, WIDTH, 4), dtype=np.uint8) for _ in range(2)]
            for hf in host_frames:
                torch.cuda.cudart().cudaHostRegister(hf.ctypes.data, hf.nbytes, 0)
        for i in range(3):
            render(i)
            assemble()
            sharded.read_frame_async(host_frames[i & 1] if rank == 0 else None)
        if rank == 0:
            rc.wait_reads()
        torch.cuda.synchronize()
        barrier()
        t = time.time()
        for i in range(args.steps):
            render(i)
            assemble()
            sharded.read_frame_async(host_frames[i & 1] if rank == 0 else None)
        if rank == 0:
            rc.wait_reads()
        torch.cuda.synchronize()
        barrier()
        e2e_s = time.time() - t
        e2e_rays = sum(rays_of(per_pose[i % len(poses)]) for i in range(args.steps))
        e2e = {"value": e2e_rays / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": 88 * world,
               "d2h_bytes_per_step": WIDTH * HEIGHT * 4, "ms_per_step": 1e3 * e2e_s / args.steps,
               "note": "no L2 flush in this leg; frame i's copy to the host overlaps the rendering of frame i+1"}
        if rank == 0:
            for hf in host_frames:
                torch.cuda.cudart().cudaHostUnregister(hf.ctypes.data)

    c5 = None
    if args.c5:
        sharded.close()
        c5 = c5_views(rc, sc, rank, world, dev, dist, torch)

    if rank == 0:
        total_rays = sum(rays_of(per_pose[i % len(poses)]) for i in range(args.steps))
        total_ms = float(step_ms.sum())
        value = total_rays / total_ms / 1e3
        peak, peak_src = measured_peak()
        # roofline of the render kernel: algorithmic bytes of the frames timed / their kernel durations
        rb = sum(alg_bytes(per_pose[p]) for p, _ in kms)
        rt = float(kt.sum()) / 1e3
        achieved = rb / rt / 1e9
        traffic = ncu_traffic()
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(meta, args, world),
            "ms_per_frame_by_pose": {str(p): float(np.mean([m for q, m in zip([k for k, _ in kms], kt) if q == p]))
                                     for p in range(len(poses))},
            "rays_per_frame_by_pose": [rays_of(c) for c in per_pose],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "peak_source": peak_src, "kernel": "render_fast_kernel" if kernel_used == 2 else "render_kernel",
                         "algorithmic_bytes_per_frame_by_pose": [alg_bytes(c) for c in per_pose],
                         "kernel_ms_mean": float(kt.mean())},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "wall_s_timed_region": wall, "kernel": {1: "generic", 2: "fast"}.get(kernel_used),
            "gather": args.gather if world > 1 else None,
            "division": {0: "glsl a*(1/b) (matches the reference shader on llvmpipe bit for bit)",
                         1: "ieee a/b (matches the reference CPU twin)"}[args.division],
        }
        out["extras"] = {}
        if c5 is not None:
            out["extras"]["c5_64_views"] = c5
        if not args.no_c1 and world == 1:
            out["extras"]["c1_640x360"] = c1_gpu(K, local)
        if not args.no_cpu and world == 1:
            cores = os.cpu_count() or 1
            rays, dt = 0, 0.0
            for rep in range(args.cpu_passes):
                for p in range(len(poses)):
                    a, b = cpu_port_sample(sc, poses[p], (0, HEIGHT), 0)
                    rays += a
                    dt += b
            out["cpu_baseline"] = {"value": rays / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": "oracle restatement of octree_fsh (both trees, shading): %d pass(es) over "
                                             "the full 1080p frame of each of the %d poses, OpenMP over rows, "
                                             "%.1f s of CPU wall time" % (args.cpu_passes, len(poses), dt)}
        emit(json.dumps(out))

    sharded.close()
    barrier()
    rc.destroy()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
