#!/usr/bin/env python
"""bench.py -- MDQE hot-path throughput on B200: MSDeformAttn fwd+bwd clips/s (+ roofline, CPU baseline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shape r50_360|r50_720|swinl_360]
                    [--dist local|uniform] [--dtype fp32|bf16] [--arm abi|module]

One *step* = the hot path of ONE training clip per GPU.  The headline configuration is the R50_ovis_360 shape
(BASELINE.json configs[1]; T=4 frames 384x640 padded, 4-level pyramid 48x80..6x10 => S=5100, 8 heads x D=32, 4 points):

  36 MSDeformAttn forward + 36 backward calls through the C ABI of libmsda_b200.so
     6 encoder layers        N=4 frames, Lq=S=5100, L=4          (transformer_enc.py:100-110)
     6 decoder spatial       N=4 frames, Lq=196,    L=4          (transformer_dec.py:340-346)
     6x4 decoder temporal    N=1 clip,   Lq=196,    L=T=4 frames of one pyramid level each (:361-395)
  + the mask-logit contraction Q=196 x K=32 x (T*96*160) forward and backward  (matcher.py:182)

run forward in layer order and backward in reverse (as autograd would), every call on its own input and
output buffers (~1.5 GB per step, so no call finds its inputs in the 126 MB L2).  Inputs are synthetic
(seed = rank), SURVEY 8(d): value ~ randn, aw = softmax(randn), loc "local" = reference point +
0.05*randn clamped to [-0.1, 1.1] (default) or "uniform" in [0,1).

Output: ONE JSON line (rank 0).  `value` is clips/s with inputs resident in HBM (CUDA-graph replay of the
step, CUDA events, max over ranks); `e2e` is the same step through the *_host C-ABI entries with pinned
host buffers (H2D of every input and D2H of every output inside the timed region); `roofline` is the
dominant kernel (encoder-shape backward) timed per launch with CUDA events on its stream; `other_configs`
carries the other BASELINE.json configurations (R50_ovis_720 training step, swinl_ytvis21 inference and training,
bf16 storage) measured the same way in the same run; `module_arm` is the step through the
`mdqe_cvpr2023_b200.MSDeformAttn` modules (Linear layers + fused sampler, torch autograd, CUDA graph);
`cpu_baseline` / `--impl reference` time oracle/torch_port.py (the reference's CPU path restated:
F.grid_sample + autograd) on the host cores.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HEADS, POINTS, QUERIES, N_LAYERS = 8, 4, 196, 6
# configs/R50_ovis_360.yaml:36,44 / R50_ovis_720.yaml:37,45 / swinl_ytvis21.yaml:33,41 (SURVEY 8 shape table)
SHAPES = {
    "r50_360": dict(name="R50_ovis_360", pyramid=[(48, 80), (24, 40), (12, 20), (6, 10)], frames=4, head_dim=32, mask_k=32,
                    mask_plane=(96, 160)),
    "r50_720": dict(name="R50_ovis_720", pyramid=[(80, 144), (40, 72), (20, 36), (10, 18)], frames=4, head_dim=32, mask_k=32,
                    mask_plane=(160, 288)),
    "swinl_360": dict(name="swinl_ytvis21", pyramid=[(48, 80), (24, 40), (12, 20), (6, 10)], frames=3, head_dim=24, mask_k=24,
                      mask_plane=(96, 160)),
}
METRIC = "msda_fwd_bwd_clips_per_s"
UNIT = "clips/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="r50_360", choices=sorted(SHAPES), help="headline workload (default: BASELINE.json configs[1])")
    ap.add_argument("--arm", default="abi", choices=["abi", "module"], help="abi: the raw C-ABI step (headline); module: only the MSDeformAttn-module step")
    ap.add_argument("--dist", default="local", choices=["local", "uniform"])
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--layers", type=int, default=N_LAYERS, help="debug: fewer layers (the result is then not a valid bench value)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-legacy", action="store_true", help="time the first-generation host path (A/B)")
    ap.add_argument("--e2e-sync-every-step", action="store_true", help="drain the host pipeline after every step instead of keeping two steps in flight (A/B)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the other BASELINE configurations and the module arm")
    ap.add_argument("--no-prezero", action="store_true", help="A/B: zero-fill grad_value inside every backward call (on the critical path) instead of on a side branch")
    ap.add_argument("--no-allreduce", action="store_true", help="debug: multi-GPU step without the gradient all-reduce (not a valid bench value)")
    ap.add_argument("--allreduce-impl", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU gradient all-reduce: 'peer' = msda_allreduce_f32 (this library's kernel over NVLink peer memory: multimem "
                         "through the switch when available, else two-shot P2P; few CTAs so that it runs beside the encoder backward), "
                         "'nccl' = torch.distributed.all_reduce")
    ap.add_argument("--allreduce-ctas", type=int, default=0,
                    help="CTAs of the peer all-reduce while it overlaps the backward (0 = auto: 6 on two GPUs, 4 on more -- through the "
                         "switch one CTA moves 70 GB/s at world 2 and 190 GB/s at world 8, profiles/r02_allreduce.md)")
    ap.add_argument("--allreduce-mode", default="graph", choices=["graph", "split"],
                    help="multi-GPU: 'graph' = ONE CUDA graph per step with the NCCL all-reduces captured on a forked branch; "
                         "'split' = round 1's 1 + n_enc graphs with the all-reduces enqueued between them")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- workload
def ref_points(torch, shapes, device="cpu"):
    pts = []
    for H, W in shapes:
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=device) + 0.5,
                                torch.arange(W, dtype=torch.float32, device=device) + 0.5, indexing="ij")
        pts.append(torch.stack([xs.reshape(-1) / W, ys.reshape(-1) / H], -1))
    return torch.cat(pts)


def build_calls(torch, shape, dist, seed, layers, device="cpu"):
    """Tensors of every MSDA call of one clip, in forward order, plus the mask operands.  device="cpu": seeded CPU tensors
    (the headline configuration: the host-buffer arm and the CPU baseline need them); a CUDA device: generated there."""
    g = torch.Generator(device=device).manual_seed(seed)
    pyr, T, D = shape["pyramid"], shape["frames"], shape["head_dim"]
    kw = dict(generator=g, device=device)
    shapes = torch.tensor(pyr, dtype=torch.long, device=device)
    sizes = shapes[:, 0] * shapes[:, 1]
    S = int(sizes.sum())
    lsi = torch.cat([sizes.new_zeros(1), sizes.cumsum(0)[:-1]])
    pix = ref_points(torch, pyr, device)

    def make_loc(ref, N, Lq, L):
        if dist == "uniform":
            return torch.rand(N, Lq, HEADS, L, POINTS, 2, **kw)
        loc = ref.view(1, Lq, 1, 1, 1, 2) + 0.05 * torch.randn(N, Lq, HEADS, L, POINTS, 2, **kw)
        return loc.clamp_(-0.1, 1.1)

    def aw_(N, Lq, L):
        return torch.softmax(torch.randn(N, Lq, HEADS, L * POINTS, **kw), -1).view(N, Lq, HEADS, L, POINTS)

    calls = []
    for _ in range(layers):                                   # encoder self-attention, queries = pixels
        calls.append(dict(kind="enc", value=torch.randn(T, S, HEADS, D, **kw), shapes=shapes,
                          lsi=lsi, loc=make_loc(pix, T, S, 4), aw=aw_(T, S, 4), go=torch.randn(T, S, HEADS * D, **kw)))
    for _ in range(layers):                                   # decoder: frame-level then clip-level cross-attention
        qref = torch.rand(QUERIES, 2, **kw)
        calls.append(dict(kind="dec_spatial", value=torch.randn(T, S, HEADS, D, **kw), shapes=shapes,
                          lsi=lsi, loc=make_loc(qref, T, QUERIES, 4), aw=aw_(T, QUERIES, 4),
                          go=torch.randn(T, QUERIES, HEADS * D, **kw)))
        value_t = torch.randn(1, T * S, HEADS, D, **kw)
        loc_t = make_loc(qref, 1, QUERIES, T)
        aw_t = aw_(1, QUERIES, T)
        go_t = torch.randn(1, QUERIES, HEADS * D, **kw)
        frame_base = torch.arange(T, device=device) * S
        for lvl in range(4):                                  # reference form: one call per pyramid level; "levels" = the T frames
            calls.append(dict(kind="dec_temporal", value=value_t, shapes=shapes[lvl].view(1, 2).expand(T, 2).contiguous(),
                              lsi=frame_base + lsi[lvl], loc=loc_t, aw=aw_t, go=go_t))
        # the same four calls as ONE grouped launch (msda_forward_grouped: G = 4 level tables, mean folded in)
        calls.append(dict(kind="dec_temporal_grouped", value=value_t, shapes=shapes.view(4, 1, 2).expand(4, T, 2).contiguous(),
                          lsi=(lsi.view(4, 1) + frame_base.view(1, T)).contiguous(), loc=loc_t, aw=aw_t, go=go_t))
    K, plane = shape["mask_k"], shape["mask_plane"]
    mask = dict(coeff=torch.tanh(torch.randn(1, QUERIES, K, **kw)), proto=torch.randn(1, K, T, *plane, **kw),
                go=torch.randn(1, QUERIES, T, *plane, **kw))
    return calls, mask


def call_dims(c):
    N, S, M, D = c["value"].shape
    _, Lq, _, L, P, _ = c["loc"].shape        # for the grouped temporal call L = T frames per level table
    return N, S, M, D, L, Lq, P


def algorithmic_bytes(c, esize, lsize):
    """fwd = value + loc + aw + out ; bwd = value, loc, aw, grad_out read + gv, gloc, gaw written.
    For temporal calls only the rows of the sampled windows count as `value`."""
    N, S, M, D, L, Lq, P = call_dims(c)
    rows = int((c["shapes"][:, 0] * c["shapes"][:, 1]).sum()) if c["kind"] == "dec_temporal" else S
    v = N * rows * M * D * esize
    smp = N * Lq * M * L * P
    o = N * Lq * M * D * esize
    return v + 3 * smp * lsize + o, 2 * v + 6 * smp * lsize + o


# -------------------------------------------------------------------------------------------- our arm
class DeviceStep:
    """All buffers of one clip resident on the GPU + pre-bound C-ABI argument lists.

    prezero=True (default): the backward's grad_value accumulators are zero-filled (msda_zero_fill) on a side stream that is
    forked at the start of the step and joined before the first backward call, so the 18 fills of 20.9 MB overlap the forward
    pass instead of sitting in front of every backward kernel; the backward calls then carry MSDA_BWD_ACC_ZEROED.  This is what
    MSDeformAttnFunction does under autograd (functions.py).  All of it is inside the timed region / the captured graph."""

    def __init__(self, torch, lib, libmod, calls, mask, device, dtype, prezero=True):
        self.torch, self.lib, self.libmod, self.device = torch, lib, libmod, device
        vt = torch.bfloat16 if dtype == "bf16" else torch.float32
        self.code = libmod.MSDA_BF16 if dtype == "bf16" else libmod.MSDA_F32
        self.mcode = self.code
        self.prezero = prezero
        self.side = torch.cuda.Stream(device) if prezero else None
        self.keep = []
        self.fwd, self.bwd, self.kinds, self.fills = [], [], [], []
        cache = {}

        def dev(t, cast=True):
            key = t.data_ptr()
            if key not in cache:
                cache[key] = (t.to(vt) if (cast and t.is_floating_point()) else t).to(device).contiguous()
            return cache[key]

        flags = libmod.BWD_ACC_ZEROED if prezero else 0
        for c in calls:
            if c["kind"] == "dec_temporal":
                continue                                      # the device step runs the grouped form instead
            N, S, M, D, L, Lq, P = call_dims(c)
            v, loc, aw, go = dev(c["value"]), dev(c["loc"]), dev(c["aw"]), dev(c["go"])
            sh, ls = dev(c["shapes"], False), dev(c["lsi"], False)
            out = torch.empty(N, Lq, M * D, dtype=vt, device=device)
            gv, gl, ga = torch.empty_like(v), torch.empty_like(loc), torch.empty_like(aw)
            ws_bytes = lib.msda_backward_workspace_bytes(self.code, N, S, M, D)
            ws = torch.empty(max(ws_bytes // 4, 1), dtype=torch.float32, device=device)
            self.keep += [v, loc, aw, go, sh, ls, out, gv, gl, ga, ws]
            grouped = c["kind"] == "dec_temporal_grouped"
            self.kinds.append(c["kind"])
            self.fwd.append((grouped, (self.code, v.data_ptr(), sh.data_ptr(), ls.data_ptr(), loc.data_ptr(), aw.data_ptr())
                             + ((N, S, M, D, 4, L, Lq, P, 0.25) if grouped else (N, S, M, D, L, Lq, P)) + (out.data_ptr(),)))
            G, Lb, scale = (4, L, 0.25) if grouped else (1, L, 1.0)
            self.bwd.append((self.code, v.data_ptr(), sh.data_ptr(), ls.data_ptr(), loc.data_ptr(), aw.data_ptr(), go.data_ptr(),
                             N, S, M, D, G, Lb, Lq, P, scale, gv.data_ptr(), gl.data_ptr(), ga.data_ptr(),
                             ws.data_ptr() if ws_bytes else None, ws_bytes, flags))
            self.fills.append((ws.data_ptr(), ws_bytes) if ws_bytes else (gv.data_ptr(), gv.numel() * gv.element_size()))
        # mask contraction: forward in the I/O dtype, backward is fp32-only (training precision of the reference)
        mc, mp = dev(mask["coeff"]), dev(mask["proto"])
        B, Q, K = mc.shape
        ncols = mp.numel() // (B * K)
        mo = torch.empty(B, Q, ncols, dtype=vt, device=device)
        self.mask_fwd = (self.mcode, self.mcode, mc.data_ptr(), mp.data_ptr(), B, Q, K, ncols, mo.data_ptr())
        c32, p32, g32 = mask["coeff"].float().to(device), mask["proto"].float().to(device), mask["go"].float().to(device)
        gc, gp = torch.empty_like(c32), torch.empty_like(p32)
        self.mask_bwd = (libmod.MSDA_F32, c32.data_ptr(), p32.data_ptr(), g32.data_ptr(), B, Q, K, ncols, gc.data_ptr(), gp.data_ptr())
        self.keep += [mc, mp, mo, c32, p32, g32, gc, gp]
        self.outputs = dict(mask=mo)
        self.n_enc = sum(1 for k in self.kinds if k == "enc")

    def _check(self, rc):
        if rc:
            from mdqe_cvpr2023_b200 import _lib
            raise RuntimeError("C-ABI call failed: " + _lib.last_error())

    def run(self, part="all"):
        """enqueue one step on the current stream (no host sync).  part "head" = forward + mask + decoder backward,
        part "tail<i>" = the i-th encoder backward (the split lets the multi-GPU run overlap the decoder-gradient all-reduce
        with the encoder backward, as DDP's buckets do); "fwd" = the inference pass (forward calls + mask logits)."""
        lib, torch = self.lib, self.torch
        cur = torch.cuda.current_stream(self.device)
        st = cur.cuda_stream
        rc = 0
        if part in ("all", "head", "fwd"):
            if self.prezero and part != "fwd":                 # fork: zero-fill every accumulator of this step on the side stream
                self.side.wait_stream(cur)
                for ptr, nbytes in self.fills:
                    rc |= lib.msda_zero_fill(self.side.cuda_stream, ptr, nbytes)
            for grouped, a in self.fwd:
                rc |= (lib.msda_forward_grouped if grouped else lib.msda_forward)(st, *a)
            rc |= lib.mask_logits_forward(st, *self.mask_fwd)
            if part == "fwd":                                  # inference: forward calls + mask logits only
                return self._check(rc)
            rc |= lib.mask_logits_backward(st, *self.mask_bwd)
            if self.prezero:
                cur.wait_stream(self.side)                     # join before the first backward
        order = list(reversed(self.bwd))                       # decoder calls first, the n_enc encoder calls last
        if part == "head":
            order = order[:len(order) - self.n_enc]
        elif part.startswith("tail"):                          # "tail<i>": the i-th encoder backward of the step (last layer first)
            i = len(order) - self.n_enc + int(part[4:])
            order = order[i:i + 1]
        for a in order:
            rc |= lib.msda_backward_grouped_flags(st, *a)
        self._check(rc)


class HostStep:
    """The same step through the host-buffer C-ABI: pinned host buffers in, pinned host buffers out.  Forward calls use the
    *_host_saved entries (device copies of the inputs stay alive like autograd's saved tensors; the backward uploads only
    grad_out), the clip-level attention goes through the grouped form (one value upload instead of four), and the mask
    head runs forward and backward -- exactly the launches of the device step.  `legacy=True` is the first-generation path
    (every call re-uploads its inputs, temporal calls per level, no mask backward) kept for A/B timing."""

    def __init__(self, torch, lib, libmod, calls, mask, device_index, dtype, legacy=False):
        self.lib, self.dev, self.legacy, self.ctypes = lib, device_index, legacy, ctypes
        vt = torch.bfloat16 if dtype == "bf16" else torch.float32
        code = libmod.MSDA_BF16 if dtype == "bf16" else libmod.MSDA_F32
        self.keep, self.fwd, self.bwd = [], [], []
        self.h2d = self.d2h = 0
        cache = {}

        def pin(t, cast=True):
            key = t.data_ptr()
            if key not in cache:
                cache[key] = (t.to(vt) if (cast and t.is_floating_point()) else t).contiguous().pin_memory()
            return cache[key]

        nb = lambda t: t.numel() * t.element_size()
        # output staging shared by all calls (max size), pinned
        big = max(calls, key=lambda c: c["go"].numel())
        bigv = max(calls, key=lambda c: c["value"].numel())
        out_h = torch.empty(big["go"].numel(), dtype=vt).pin_memory()
        gv_h = torch.empty(bigv["value"].numel(), dtype=vt).pin_memory()
        gl_h = torch.empty(big["loc"].numel(), dtype=vt).pin_memory()
        ga_h = torch.empty(big["aw"].numel(), dtype=vt).pin_memory()
        self.keep += [out_h, gv_h, gl_h, ga_h]
        skip = "dec_temporal_grouped" if legacy else "dec_temporal"
        for c in calls:
            if c["kind"] == skip:
                continue
            N, S, M, D, L, Lq, P = call_dims(c)
            v, loc, aw, go = pin(c["value"]), pin(c["loc"]), pin(c["aw"]), pin(c["go"])
            sh, ls = pin(c["shapes"], False), pin(c["lsi"], False)
            self.keep += [v, loc, aw, go, sh, ls]
            small = nb(sh) + nb(ls)
            if legacy:
                dims = (N, S, M, D, L, Lq, P)
                self.fwd.append((device_index, code, v.data_ptr(), sh.data_ptr(), ls.data_ptr(), loc.data_ptr(), aw.data_ptr())
                                + dims + (out_h.data_ptr(),))
                self.bwd.append((device_index, code, v.data_ptr(), sh.data_ptr(), ls.data_ptr(), loc.data_ptr(), aw.data_ptr(),
                                 go.data_ptr()) + dims + (gv_h.data_ptr(), gl_h.data_ptr(), ga_h.data_ptr()))
                self.h2d += 2 * (nb(v) + nb(loc) + nb(aw) + small) + nb(go)
            else:
                G, scale = (4, 0.25) if c["kind"] == "dec_temporal_grouped" else (1, 1.0)
                self.fwd.append((device_index, code, v.data_ptr(), sh.data_ptr(), ls.data_ptr(), loc.data_ptr(), aw.data_ptr(),
                                 N, S, M, D, G, L, Lq, P, scale, out_h.data_ptr()))
                self.bwd.append((go.data_ptr(), gv_h.data_ptr(), gl_h.data_ptr(), ga_h.data_ptr()))
                self.h2d += nb(v) + nb(loc) + nb(aw) + small + nb(go)
            self.d2h += nb(go) + nb(v) + nb(loc) + nb(aw)
        mc, mp = pin(mask["coeff"]), pin(mask["proto"])
        B, Q, K = mc.shape
        ncols = mp.numel() // (B * K)
        mo = torch.empty(B * Q * ncols, dtype=vt).pin_memory()
        self.keep += [mc, mp, mo]
        self.mask_fwd = (device_index, code, code, mc.data_ptr(), mp.data_ptr(), B, Q, K, ncols, mo.data_ptr())
        self.h2d += nb(mc) + nb(mp)
        self.d2h += nb(mo)
        self.mask_bwd = None
        if not legacy and dtype != "bf16":                    # the mask backward is fp32-only (the reference's training precision)
            mg = pin(mask["go"])
            gc_h, gp_h = torch.empty_like(mc).pin_memory(), torch.empty_like(mp).pin_memory()
            self.keep += [mg, gc_h, gp_h]
            self.mask_bwd = (mg.data_ptr(), gc_h.data_ptr(), gp_h.data_ptr())
            self.h2d += nb(mg)
            self.d2h += nb(gc_h) + nb(gp_h)

    def run(self):
        lib, rc = self.lib, 0
        if self.legacy:
            for a in self.fwd:
                rc |= lib.msda_forward_host(*a)
            rc |= lib.mask_logits_forward_host(*self.mask_fwd)
            for a in reversed(self.bwd):
                rc |= lib.msda_backward_host(*a)
        else:
            c_i64 = self.ctypes.c_int64
            handles = []
            for a in self.fwd:
                h = c_i64(0)
                rc |= lib.msda_forward_host_saved(*a, self.ctypes.byref(h))
                handles.append(h.value)
            if self.mask_bwd is not None:
                h = c_i64(0)
                rc |= lib.mask_logits_forward_host_saved(*self.mask_fwd, self.ctypes.byref(h))
                rc |= lib.mask_logits_backward_host_saved(h.value, *self.mask_bwd)
            else:
                rc |= lib.mask_logits_forward_host(*self.mask_fwd)
            for h, a in zip(reversed(handles), reversed(self.bwd)):
                rc |= lib.msda_backward_host_saved(h, *a)
        if rc:
            from mdqe_cvpr2023_b200 import _lib
            raise RuntimeError("C-ABI host call failed: " + _lib.last_error())


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = the upper half of the samples by power draw
        order = sorted(range(len(sm)), key=lambda i: power[i])
        load = [sm[i] for i in order[len(order) // 2:]]
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_sample(torch, calls, mask, reps, warm):
    """Reference CPU path (torch port) on one layer's calls + the mask contraction; returns seconds
    (layer part, mask part), averaged over `reps`."""
    from oracle import torch_port as TP
    layer = [next(c for c in calls if c["kind"] == "enc"), next(c for c in calls if c["kind"] == "dec_spatial")] + \
            [c for c in calls if c["kind"] == "dec_temporal"][:4]
    t_layer, t_mask = [], []
    for it in range(warm + reps):
        t0 = time.perf_counter()
        for c in layer:
            TP.msda_fwd_bwd_torch(c["value"], c["shapes"], c["loc"], c["aw"], c["go"], c["lsi"])
        t1 = time.perf_counter()
        coeff = mask["coeff"].clone().requires_grad_(True)
        proto = mask["proto"].clone().requires_grad_(True)
        TP.mask_logits_torch(coeff, proto).backward(mask["go"])
        t2 = time.perf_counter()
        if it >= warm:
            t_layer.append(t1 - t0)
            t_mask.append(t2 - t1)
    return sum(t_layer) / len(t_layer), sum(t_mask) / len(t_mask)


def workload_text(shape):
    T, pyr, D, K = shape["frames"], shape["pyramid"], shape["head_dim"], shape["mask_k"]
    S = sum(h * w for h, w in pyr)
    ncols = T * shape["mask_plane"][0] * shape["mask_plane"][1]
    return (f"{shape['name']} clip hot path: 36 MSDeformAttn fwd+bwd (6 enc N={T} S=Lq={S}, 6 dec-spatial Lq={QUERIES}, "
            f"24 dec-temporal L=T={T} run as 6 grouped launches) + mask contraction Q={QUERIES} K={K} N={ncols} fwd+bwd")


def config_dict(args, shape):
    """identical for both arms (the driver compares the two `config` objects)"""
    return {"workload": workload_text(shape), "clips_per_gpu_per_step": 1, "frames": shape["frames"], "pyramid": shape["pyramid"],
            "heads": HEADS, "head_dim": shape["head_dim"], "points": POINTS, "queries": QUERIES, "layers": args.layers,
            "loc_dist": args.dist,
            "l2_policy": "inputs larger than L2 (every call has its own buffers, ~1.5 GB touched per step)",
            "parallelism": f"clip-sharded x{args.gpus}"}


# --------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's CPU implementation of the path (oracle/torch_port.py: F.grid_sample + autograd, the restatement of
    ms_deform_attn_core_pytorch, ms_deform_attn_func.py:45-65) on all host cores.  One timed step = a bounded sample of the
    clip: ONE of the six identical layers (1 encoder + 1 decoder-spatial + 4 decoder-temporal MSDA fwd+bwd) + the mask
    contraction fwd+bwd; `ms_per_step` is that measured sample, `value` the clips/s it implies (clip = 6 layers + mask)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    shape = SHAPES[args.shape]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    calls, mask = build_calls(torch, shape, args.dist, 0, 1)
    t_layer, t_mask = [], []
    for it in range(args.warmup + args.steps):
        a, b = cpu_sample(torch, calls, mask, 1, 0)
        if it >= args.warmup:
            t_layer.append(a)
            t_mask.append(b)
    tl, tm = sum(t_layer) / len(t_layer), sum(t_mask) / len(t_mask)
    clip_s = args.layers * tl + tm
    value = 1.0 / clip_s
    sample = (f"each timed step = 1 of the {args.layers} identical layers (1 enc + 1 dec-spatial + 4 dec-temporal MSDA fwd+bwd) + mask fwd+bwd "
              f"= ms_per_step; clip time = {args.layers} * layer + mask (clip_ms), value = 1 / clip time")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": (tl + tm) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, shape),
            "step_is": "bounded sample: 1 layer + mask (see cpu_baseline.sample)", "clip_ms": clip_s * 1e3,
            "layer_ms": tl * 1e3, "mask_ms": tm * 1e3, "device": "cpu", "threads": torch.get_num_threads(),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- helpers
def capture(torch, fn):
    """warm `fn` on a side stream, then capture it into a CUDA graph"""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def time_replays(torch, graph, steps, warm=3):
    for _ in range(warm):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def enc_kernel_times(torch, libmod, step, enc_pairs, reps=5):
    """per-launch CUDA-event times of the encoder-shape sampling kernels (eager pass, events right around the kernels)"""
    libmod.set_option("profile", 1)
    for _ in range(reps):
        step.run()
    torch.cuda.synchronize()
    bwd_ms, bwd_n = libmod.profile_read(libmod.PROF_MSDA_BWD, enc_pairs)
    fwd_ms, fwd_n = libmod.profile_read(libmod.PROF_MSDA_FWD, enc_pairs)
    libmod.set_option("profile", 0)
    return (fwd_ms / fwd_n * 1e3 if fwd_n else None), (bwd_ms / bwd_n * 1e3 if bwd_n else None)


def measure_other_config(torch, lib, libmod, key, dtype, dist, device, steps, hbm_peak):
    """One of the other BASELINE.json configurations, measured like the headline: CUDA-graph replay of the clip's training step
    and of its inference pass, inputs resident (generated on the device), plus the encoder-shape kernels per launch."""
    shape = SHAPES[key]
    calls, mask = build_calls(torch, shape, dist, 1, N_LAYERS, device=device)
    step = DeviceStep(torch, lib, libmod, calls, mask, device, dtype)
    for _ in range(2):
        step.run()
    torch.cuda.synchronize()
    train_ms = time_replays(torch, capture(torch, step.run), steps)
    fwd_ms = time_replays(torch, capture(torch, lambda: step.run("fwd")), steps)
    enc = next(c for c in calls if c["kind"] == "enc")
    esize = 2 if dtype == "bf16" else 4
    fwd_b, bwd_b = algorithmic_bytes(enc, esize, esize)
    enc_pairs = shape["frames"] * sum(h * w for h, w in shape["pyramid"]) * HEADS
    f_us, b_us = enc_kernel_times(torch, libmod, step, enc_pairs)
    res = {"workload": workload_text(shape), "dtype": "bf16" if dtype == "bf16" else "f32", "loc_dist": dist,
           "train_step_ms": train_ms, "train_clips_per_s": 1e3 / train_ms,
           "inference_ms": fwd_ms, "inference_clips_per_s": 1e3 / fwd_ms,
           "enc_fwd_us": f_us, "enc_bwd_us": b_us, "enc_fwd_bytes": fwd_b, "enc_bwd_bytes": bwd_b,
           "enc_fwd_bwd_frac_of_hbm_peak": ((fwd_b + bwd_b) / ((f_us + b_us) * 1e-6) / 1e9 / hbm_peak) if (f_us and b_us) else None,
           "timing": f"CUDA-graph replay x{steps}, CUDA events; enc_* = per-launch CUDA events around the encoder-shape kernels"}
    del step, calls, mask
    torch.cuda.empty_cache()
    return res


def module_arm(torch, shape, device, steps, dist="local", world=1):
    """The same clip through `mdqe_cvpr2023_b200.MSDeformAttn` (what train_net.py would run): 6 encoder self-attention modules on
    the [T, S, 256] pyramid, then 6 x (frame-level + clip-level) decoder cross-attention modules, residual connections between
    them, forward + backward with every parameter gradient; Linear layers as 3xTF32 tensor-core GEMMs, softmax / location
    arithmetic inside the sampler, grad_value accumulators zero-filled on a side stream.  Timed as one replayed CUDA graph."""
    from mdqe_cvpr2023_b200 import MSDeformAttn
    T, pyr, D = shape["frames"], shape["pyramid"], shape["head_dim"]
    C = HEADS * D
    S = sum(h * w for h, w in pyr)
    torch.manual_seed(0)
    enc = [MSDeformAttn(C, 4, HEADS, POINTS, pred_offsets=True, mode="spatial").to(device) for _ in range(N_LAYERS)]
    dec_f = [MSDeformAttn(C, 4, HEADS, POINTS, pred_offsets=False, mode="spatial").to(device) for _ in range(N_LAYERS)]
    dec_c = [MSDeformAttn(C, 4, HEADS, POINTS, n_frames=T, pred_offsets=False, mode="temporal").to(device) for _ in range(N_LAYERS)]
    params = [p for m in enc + dec_f + dec_c for p in m.parameters()]
    g = torch.Generator(device=device).manual_seed(0)
    for m in enc:                                              # trained-like offsets: ~0.05 of the image around the reference point
        m.sampling_offsets.weight.data.normal_(0, 0.4 / C ** 0.5, generator=g)
    shapes = torch.tensor(pyr, device=device)
    src = torch.randn(T, S, C, device=device, generator=g).requires_grad_(True)
    pix = ref_points(torch, pyr, device)
    enc_ref = torch.cat([pix, torch.full_like(pix, 0.1)], -1).unsqueeze(0).expand(T, S, 4).contiguous()
    q_f = torch.randn(T, QUERIES, C, device=device, generator=g).requires_grad_(True)
    q_c = torch.randn(1, QUERIES, C, device=device, generator=g).requires_grad_(True)
    box = torch.cat([torch.rand(QUERIES, 2, device=device, generator=g), torch.full((QUERIES, 2), 0.1, device=device)], -1)
    ref_f, ref_c = box.unsqueeze(0).expand(T, QUERIES, 4).contiguous(), box.unsqueeze(0).contiguous()

    def forward():
        x = src
        for m in enc:
            x = x + m(x, enc_ref, x, shapes, None)
        qf, qc = q_f, q_c
        mem_c = x.view(1, T, S, C)
        for mf, mc in zip(dec_f, dec_c):
            qf = qf + mf(qf, ref_f, x, shapes, None)
            qc = qc + mc(qc, ref_c, mem_c, shapes, None)
        return x, qf, qc

    def train_step():
        x, qf, qc = forward()
        return torch.autograd.grad(x.sum() + qf.sum() + qc.sum(), [src, q_f, q_c] + params)

    peer = None
    if world > 1:
        try:                                                   # this library's all-reduce over NVLink peer memory; NCCL if unavailable
            from mdqe_cvpr2023_b200.collectives import PeerAllReduce
            peer = PeerAllReduce(sum(p.numel() for p in params), device, n_ctas=16)
        except Exception:  # noqa: BLE001
            peer = None

    def ddp_step():
        # clip-sharded data parallel training (train_net.py:256-271): every rank runs its own clip, then the parameter gradients
        # THIS step produced are averaged over the ranks -- by msda_allreduce_f32 (collectives.allreduce_mean_gradients_peer) or,
        # where symmetric memory is unavailable, by the bucketed NCCL all-reduce of sharding.py -- all captured in one graph
        from mdqe_cvpr2023_b200.collectives import allreduce_mean_gradients_peer
        from mdqe_cvpr2023_b200.sharding import allreduce_mean_gradients
        for p in params:
            p.grad = None
        x, qf, qc = forward()
        (x.sum() + qf.sum() + qc.sum()).backward()
        if peer is not None:
            allreduce_mean_gradients_peer(params, peer)
        else:
            allreduce_mean_gradients(params)

    def infer_step():
        with torch.no_grad():
            return forward()

    ddp = None
    if world > 1:
        import torch.distributed as tdist
        for p in params:                                       # replicas start from the same weights (rank 0's), inputs differ per rank
            tdist.broadcast(p.data, 0)
        g_ddp = capture(torch, ddp_step)
        ddp_ms = time_replays(torch, g_ddp, steps)
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        sums = [torch.zeros(2, device=device, dtype=torch.float64) for _ in range(world)]
        tdist.all_gather(sums, torch.stack([flat.double().sum(), flat.double().abs().sum()]))
        same = all(bool((t == sums[0]).all()) for t in sums)
        ddp = {"train_step_ms": ddp_ms, "aggregate_clips_per_s": world * 1e3 / ddp_ms, "grad_elements_reduced": int(flat.numel()),
               "grad_abs_sum": float(sums[0][1]), "grads_identical_on_all_ranks": same,
               "allreduce": ("msda_allreduce_f32 (%s over NVLink peer memory)" % peer.algo) if peer is not None else "NCCL (sharding.allreduce_mean_gradients)",
               "what": "the same module step on every rank (own clip), then the parameter gradients it produced averaged over the ranks "
                       "(one concatenation, one all-reduce kernel, one multi-tensor copy back), all inside one captured CUDA graph"}
        if peer is not None:
            torch.cuda.synchronize()
            peer.check()
        del g_ddp
    train_ms = time_replays(torch, capture(torch, train_step), steps)
    infer_ms = time_replays(torch, capture(torch, infer_step), steps)
    # the same inference pass with value kept only as the sampler's paired-corner bf16 layout where that pays: the 6 encoder modules
    # (value_proj GEMM epilogue -> packed sampler with the fused prologue; value rounded to bf16, everything else fp32); the
    # decoder modules' 196-query calls stay fp32 (MSDeformAttn._packed_inference_ok)
    packed = None
    if D == 32:
        try:
            want = [t.clone() for t in infer_step()]
            for m in enc + dec_f:
                m.value_storage = "bf16_packed"
            got = infer_step()
            err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(got, want))
            packed_ms = time_replays(torch, capture(torch, infer_step), steps)
            packed = {"inference_ms": packed_ms, "inference_clips_per_s": 1e3 / packed_ms, "max_normalised_error_vs_fp32": err,
                      "what": "value_storage = 'bf16_packed': the 6 encoder modules run tc_linear_forward_packed + msda_fused_forward_packed_joint (the decoder's 196-query calls stay fp32: the packed epilogue does not pay there)"}
        except Exception as e:  # noqa: BLE001
            packed = {"error": repr(e)[:300]}
        finally:
            for m in enc + dec_f:
                m.value_storage = "fp32"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        train_step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        train_step()
    e1.record()
    torch.cuda.synchronize()
    return {"what": "18 MSDeformAttn modules of one clip (6 encoder, 6 frame-level + 6 clip-level decoder) forward+backward incl. all "
                    "parameter gradients, through mdqe_cvpr2023_b200.MSDeformAttn (tc_linear + fused prologue + grouped temporal launch)",
            "shape": shape["name"], "train_step_ms": train_ms, "train_clips_per_s": 1e3 / train_ms, "inference_ms": infer_ms,
            "inference_clips_per_s": 1e3 / infer_ms, "eager_train_step_ms": e0.elapsed_time(e1) / 5,
            "inference_bf16_packed": packed,
            "parameters": sum(p.numel() for p in params), "timing": f"CUDA-graph replay x{steps}, CUDA events", "ddp": ddp}


# ----------------------------------------------------------------------------------------------- main
def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    os.environ.setdefault("NCCL_DEBUG", "WARN")       # keep NCCL's version banner off stdout (one JSON line only)
    # Gradient buckets overlap with the encoder backward, whose kernels are bound by per-SM pipes: the ring (LL128, 32 channels)
    # disturbs them less than the NVLink-SHARP (NVLS) all-reduce NCCL picks by default on an NVSwitch box (8 GPUs: 2.10 vs 2.18 ms/step)
    os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
    import torch
    import torch.distributed as dist
    from mdqe_cvpr2023_b200 import _lib as libmod
    lib = libmod.load()
    shape = SHAPES[args.shape]
    T, PYRAMID, HEAD_DIM, MASK_K, MASK_PLANE = shape["frames"], shape["pyramid"], shape["head_dim"], shape["mask_k"], shape["mask_plane"]

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator comes up; the contract is ONE JSON
        # line there, so park fd 1 on stderr while the process group is created and exercised once.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            warm = torch.zeros(1, device=device)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json"
    else:
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    if args.arm == "module":                                   # only the module-level step
        res = module_arm(torch, shape, device, args.steps, args.dist, world)
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": world * res["train_clips_per_s"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "warmup": 3, "ms_per_step": res["train_step_ms"], "higher_is_better": True, "scaling": "weak",
                              "vs_baseline": None, "dtype": "f32", "data": "synthetic", "arm": "module", "config": config_dict(args, shape),
                              "module_arm": res}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    calls, mask = build_calls(torch, shape, args.dist, rank, args.layers)
    step = DeviceStep(torch, lib, libmod, calls, mask, device, args.dtype, prezero=not args.no_prezero)
    # gradient all-reduce of the enc+dec parameters (decoder 14.94 M, encoder 4.54 M fp32) -- the only
    # cross-GPU step of clip-sharded DDP training.  Buckets like DDP's, in the order the gradients become ready: the decoder
    # bucket is reduced on NCCL's stream while the encoder backward runs, then one bucket per encoder layer as soon as that
    # layer's backward has been enqueued; only the last layer's 3 MB are reduced after the backward has ended.
    n_enc_layers = step.n_enc
    DEC_ELEMS = 14_940_000
    enc_elems = (4_540_000 // max(n_enc_layers, 1) + 3) // 4 * 4
    peer_ar, peer_note = None, None
    dec_buf = enc_bufs = None
    if args.allreduce_ctas <= 0:
        args.allreduce_ctas = 6 if world == 2 else 4
    if world > 1 and args.allreduce_impl == "peer":
        try:
            from mdqe_cvpr2023_b200.collectives import PeerAllReduce
            peer_ar = PeerAllReduce(DEC_ELEMS + n_enc_layers * enc_elems, device, n_ctas=args.allreduce_ctas)
        except Exception as e:  # noqa: BLE001 -- no symmetric memory on this box / torch build: NCCL does the reduction
            peer_note = "peer all-reduce unavailable (%s); NCCL used" % repr(e)[:200]
    if world > 1 and peer_ar is None:
        dec_buf = torch.zeros(DEC_ELEMS, device=device)
        enc_bufs = [torch.zeros(enc_elems, device=device) for _ in range(n_enc_layers)]

    def reduce_bucket(i, last=False):
        """bucket -1 = decoder parameters, i >= 0 = encoder layer i (gradient-ready order); enqueues on the current stream"""
        if peer_ar is not None:
            off, n = (0, DEC_ELEMS) if i < 0 else (DEC_ELEMS + i * enc_elems, enc_elems)
            # the last bucket has nothing left to hide behind: give it enough CTAs to finish quickly
            peer_ar.all_reduce_(off, n, mean=True, n_ctas=16 if last else args.allreduce_ctas)
        else:
            dist.all_reduce(dec_buf if i < 0 else enc_bufs[i])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up (eager), then capture one step into a CUDA graph
    for _ in range(max(args.warmup, 3)):
        step.run()
        if world > 1:
            for i in range(-1, n_enc_layers):
                reduce_bucket(i, last=(i == n_enc_layers - 1))
    torch.cuda.synchronize()
    graphs = None
    # high priority: the few CTAs of a bucket's all-reduce are dispatched as soon as an SM has room, not behind the 3400 CTAs of the backward
    nccl_side = torch.cuda.Stream(device, priority=-1) if world > 1 else None

    def whole_step(with_allreduce=True, st=None):
        """One multi-GPU step for capture: the decoder-gradient bucket is all-reduced on a forked branch as soon as the decoder
        backward has been enqueued, one encoder bucket after each encoder layer's backward; the branch joins at the end.
        st: the DeviceStep to run (default: the headline shape's)."""
        st = step if st is None else st
        cur = torch.cuda.current_stream(device)
        st.run("head")
        if with_allreduce:
            nccl_side.wait_stream(cur)
            with torch.cuda.stream(nccl_side):
                reduce_bucket(-1)
        for i in range(n_enc_layers):
            st.run(f"tail{i}")
            if with_allreduce:
                nccl_side.wait_stream(cur)
                with torch.cuda.stream(nccl_side):
                    reduce_bucket(i, last=(i == n_enc_layers - 1))
        if with_allreduce:
            cur.wait_stream(nccl_side)

    one_graph = world > 1 and args.allreduce_mode == "graph" and not args.no_graph
    graph_note = None
    if one_graph:
        try:
            graphs = [capture(torch, lambda: whole_step(not args.no_allreduce))]
        except Exception as e:  # noqa: BLE001 -- NCCL builds that cannot be captured fall back to the split form
            graph_note = "whole-step capture failed (%s); fell back to split graphs" % repr(e)[:200]
            one_graph, graphs = False, None
            torch.cuda.synchronize()
    if graphs is None and not args.no_graph:
        graphs = [capture(torch, (lambda p=part: step.run(p)))
                  for part in (("all",) if world == 1 else ("head",) + tuple(f"tail{i}" for i in range(n_enc_layers)))]
    if graphs is not None:
        for _ in range(2):
            for g in graphs:
                g.replay()
        torch.cuda.synchronize()

    def one_step():
        if world == 1 or one_graph:
            if graphs is not None:
                graphs[0].replay()
            else:
                step.run()
            return
        if graphs is not None:
            graphs[0].replay()
        else:
            step.run("head")
        cur = torch.cuda.current_stream(device)
        if not args.no_allreduce:                               # overlaps with the encoder backward below
            nccl_side.wait_stream(cur)
            with torch.cuda.stream(nccl_side):
                reduce_bucket(-1)
        for i in range(n_enc_layers):
            if graphs is not None:
                graphs[1 + i].replay()
            else:
                step.run(f"tail{i}")
            if not args.no_allreduce:
                nccl_side.wait_stream(cur)
                with torch.cuda.stream(nccl_side):
                    reduce_bucket(i, last=(i == n_enc_layers - 1))
        cur.wait_stream(nccl_side)

    # ---- timed region: exactly K steps, CUDA events, max over ranks
    libmod.launch_count_reset()
    sync_all()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    sync_all()
    if peer_ar is not None:
        peer_ar.check()                                # a peer that never reached a barrier would have been reported here
    clocks = sampler.stop() if sampler else None
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    launches_per_step = len(step.fwd) + len(step.bwd) + 2 + (len(step.bwd) if args.dtype == "bf16" else 0)
    gpu_launches = launches_per_step * args.steps      # graph replays re-launch the captured kernels
    value = world * 1.0 / (ms_step / 1e3)

    # ---- per-launch kernel timing (eager pass over the same buffers, events right around the kernels)
    esize = 2 if args.dtype == "bf16" else 4
    enc = next(c for c in calls if c["kind"] == "enc")
    enc_pairs = T * sum(h * w for h, w in PYRAMID) * HEADS
    libmod.set_option("profile", 1)
    for _ in range(min(args.steps, 10)):
        step.run()
    torch.cuda.synchronize()
    bwd_ms, bwd_n = libmod.profile_read(libmod.PROF_MSDA_BWD, enc_pairs)
    fwd_ms, fwd_n = libmod.profile_read(libmod.PROF_MSDA_FWD, enc_pairs)
    mfw_ms, mfw_n = libmod.profile_read(libmod.PROF_MASK_FWD, 0)
    mbw_ms, mbw_n = libmod.profile_read(libmod.PROF_MASK_BWD, 0)
    libmod.set_option("profile", 0)
    fwd_bytes, bwd_bytes = algorithmic_bytes(enc, esize, esize)

    traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic_db = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    main_shape = args.shape == "r50_360"                       # the committed ncu captures are of this shape

    def roof(nbytes, ms, n, name):
        if not n:
            return None
        us = ms / n * 1e3
        ach = nbytes / (us * 1e-6) / 1e9
        # DRAM bytes of one launch from the committed `ncu --set full` capture of the same kernel and shape (fp32)
        traffic = traffic_db.get(name.split("<")[0].split(":")[0]) if (args.dtype == "fp32" and main_shape) else None
        return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_db.get("source") if traffic else None, "avg_launch_us": us, "launches_timed": n, "algorithmic_bytes": nbytes, "peak_source": peak_src}

    vt_name = "bf16" if args.dtype == "bf16" else "float"
    enc_desc = f"(encoder shape N={T},S=Lq={sum(h * w for h, w in PYRAMID)})"
    roofline = roof(bwd_bytes, bwd_ms, bwd_n, f"msda_bwd_fast2_kernel<{vt_name},{HEAD_DIM},16> {enc_desc}")
    # What actually binds the two sampling kernels: the SM's memory pipeline -- every gathered corner row, every
    # vector reduction, every shared-memory access and every shuffle is a wavefront on the L1 data pipe (one per clock).
    # Wavefronts per launch come from the committed ncu capture of the same kernel and shape (profiles/ncu_traffic.json); the
    # time is the live per-launch figure above and the clock the one sampled during the timed region.  The backward's second
    # ceiling, L2's fp32 reduction rate (tools/red_microbench.cu), is reported beside it.
    def l1_pipe(roof_entry, key):
        if not (roof_entry and args.dtype == "fp32" and args.dist == "local" and main_shape and traffic_db.get(key)):
            return None
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        peak = 148 * mhz * 1e6 / 1e9                                  # G wavefronts / s
        ach = traffic_db[key] / (roof_entry["avg_launch_us"] * 1e-6) / 1e9
        return {"kernel": roof_entry["kernel"], "bound": "l1_data_pipe_wavefronts", "achieved": ach, "peak": peak, "unit": "Gwavefront/s",
                "frac": ach / peak, "wavefronts_per_launch": traffic_db[key], "source": traffic_db.get("l1_wavefronts_source")}

    binding = l1_pipe(roofline, "msda_bwd_fast2_kernel_l1_wavefronts")
    l2_reduction = None
    if roofline and args.dtype == "fp32" and args.dist == "local" and main_shape and traffic_db.get("msda_bwd_fast2_kernel_l2_red_sectors"):
        red_bytes = 32.0 * traffic_db["msda_bwd_fast2_kernel_l2_red_sectors"]
        ach = red_bytes / (roofline["avg_launch_us"] * 1e-6) / 1e9
        l2_reduction = {"kernel": roofline["kernel"], "bound": "l2_fp32_reductions", "achieved": ach, "peak": traffic_db["l2_red_peak_gbs"],
                        "unit": "GB/s", "frac": ach / traffic_db["l2_red_peak_gbs"], "reduction_bytes_per_launch": red_bytes,
                        "peak_source": traffic_db["l2_red_peak_source"]}
    roofline_fwd = roof(fwd_bytes, fwd_ms, fwd_n, f"msda_fwd_fast2_kernel<{vt_name},{HEAD_DIM},16> {enc_desc}")
    mB = QUERIES * MASK_K + MASK_K * T * MASK_PLANE[0] * MASK_PLANE[1]
    mO = QUERIES * T * MASK_PLANE[0] * MASK_PLANE[1]
    binding_fwd = l1_pipe(roofline_fwd, "msda_fwd_fast2_kernel_l1_wavefronts")
    roofline_mask = roof(mB * esize + mO * esize, mfw_ms, mfw_n, "mask_fwd_tc2_kernel<bf16> (tcgen05)" if args.dtype == "bf16" else "mask_fwd_tc4_kernel<float,false> (tcgen05, 3xTF32)")
    if roofline_mask:
        flops = 2.0 * QUERIES * MASK_K * T * MASK_PLANE[0] * MASK_PLANE[1]
        roofline_mask["tflops"] = flops / (roofline_mask["avg_launch_us"] * 1e-6) / 1e12
    roofline_mask_bwd = roof((mB + mO + mB) * 4, mbw_ms, mbw_n, "mask_backward_tc: mask_grad_coeff_tc_kernel + mask_fwd_tc4_kernel<float,true> (tcgen05, 3xTF32)")

    # ---- forward only (BASELINE configs[1]: the inference pass of the same clip -- 36 MSDeformAttn forward calls + mask logits)
    fwd_only = None
    if world == 1 and not args.no_graph:
        f_ms = time_replays(torch, capture(torch, lambda: step.run("fwd")), args.steps)
        fwd_only = {"ms_per_step": f_ms, "clips_per_s": 1e3 / f_ms, "launches_per_step": len(step.fwd) + 1,
                    "what": "forward calls of the same clip + mask logits (inference pass), CUDA-graph replay, inputs resident"}

    # ---- eager (no graph) step time, for reference
    sync_all()
    libmod.launch_count_reset()
    e0.record()
    for _ in range(min(args.steps, 10)):
        step.run()
    e1.record()
    torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1) / min(args.steps, 10)
    counted_per_step = libmod.launch_count() / min(args.steps, 10)

    # ---- the step with the zero-fills on the critical path (what round 1 measured), for comparison
    inline_fill = None
    if world == 1 and not args.no_graph and not args.no_prezero and not args.no_other_configs:
        step.prezero = False
        saved_bwd = step.bwd
        step.bwd = [a[:-1] + (0,) for a in saved_bwd]
        inline_fill = {"ms_per_step": time_replays(torch, capture(torch, step.run), args.steps),
                       "what": "same step with grad_value zero-filled inside every backward call (cudaMemsetAsync in front of the kernel)"}
        step.bwd, step.prezero = saved_bwd, True

    # ---- end to end through the host-buffer C ABI (H2D + kernels + D2H every call), wall clock
    e2e = None
    if not args.no_e2e:
        host = HostStep(torch, lib, libmod, calls, mask, local_rank, args.dtype, legacy=args.e2e_legacy)
        libmod.set_option("host_async", 1)           # calls enqueue on the library's H2D / compute / D2H streams ...

        def fence():
            t = ctypes.c_int64(0)
            libmod.check(lib.msda_host_fence(ctypes.byref(t)), "msda_host_fence")
            return t.value

        def e2e_steps(n):
            # ... every step ends with a fence, and a step's ticket is waited for (all its results in host memory) while the NEXT
            # step is already enqueued: at most two steps in flight, like a trainer that prefetches the next clip.  Step i+1's
            # forward uploads then run under step i's backward downloads (PCIe full duplex across the step boundary).
            prev = None
            for _ in range(n):
                host.run()
                t = fence()
                if args.e2e_sync_every_step:
                    libmod.check(lib.msda_host_sync(), "msda_host_sync")
                elif prev is not None:
                    libmod.check(lib.msda_host_wait(prev), "msda_host_wait")
                prev = t
            libmod.check(lib.msda_host_wait(prev), "msda_host_wait")

        e2e_steps(3)                                 # warm-up: arena allocation, saved-block pool for two steps in flight, page faults
        libmod.check(lib.msda_host_sync(), "msda_host_sync")
        n_e2e = max(1, min(args.steps, 10))
        sync_all()
        t0 = time.perf_counter()
        e2e_steps(n_e2e)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        libmod.set_option("host_async", 0)
        if world > 1:
            t = torch.tensor([dt], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        gbs = world * (host.h2d + host.d2h) * n_e2e / dt / 1e9
        e2e = {"value": world * n_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": host.h2d, "d2h_bytes_per_step": host.d2h,
               "steps": n_e2e, "ms_per_step": dt / n_e2e * 1e3, "host_link_gbs_all_ranks": gbs,
               "limit": "host<->device link: the step moves h2d+d2h bytes per rank through PCIe.  Measured on this pool's 8-GPU box with all ranks copying "
                        "at once (tools/hostlink_probe.py, profiles/r02_hostlink.md): 78 GB/s full duplex for one rank alone, but only 154 GB/s in total for 8 "
                        "ranks (16-23 GB/s per rank), i.e. >= 79 ms per step at 8 GPUs whatever the kernels do -- the host side of the box, not NVLink or the GPU",
               "timing": ("wall clock; *_host C-ABI calls in host_async mode (3-stream pipeline), msda_host_sync() at the end of every step" if args.e2e_sync_every_step else
                          "wall clock over all steps; *_host C-ABI calls in host_async mode (3-stream pipeline); every step ends with msda_host_fence() and is "
                          "waited for (msda_host_wait: all its results in host memory) while the next step is enqueued -- at most two steps in flight"),
               "note": ("first-generation host path: every call re-uploads its inputs, per-level temporal calls, no mask backward" if args.e2e_legacy else
                        "same launches as the device step; forward inputs stay on the device for the backward (*_host_saved entries)")}
        lib.msda_host_arena_release()
        del host

    # ---- multi-GPU: the same captured step without the collectives (what the all-reduce costs), and the module-level step with
    # the parameter gradients it produces averaged over the ranks
    no_coll_ms, modarm_ddp = None, None
    if world > 1 and one_graph and not args.no_allreduce:
        g_nc = capture(torch, lambda: whole_step(False))
        sync_all()
        no_coll_ms = time_replays(torch, g_nc, args.steps)
        t = torch.tensor([no_coll_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        no_coll_ms = float(t.item())
        del g_nc
    if world > 1 and not args.no_other_configs and not args.no_graph:
        try:
            modarm_ddp = module_arm(torch, shape, device, max(3, min(args.steps, 10)), args.dist, world)
        except Exception as e:  # noqa: BLE001
            modarm_ddp = {"error": repr(e)[:300]}

    # ---- multi-GPU runs: BASELINE.json configs 3 and 4 at this GPU count: the R50_ovis_720 DDP training step with the
    # same gradient all-reduce, and the Swin-L clip-sharded inference pass (no collective); max over ranks, aggregate over all GPUs
    other_multi = None
    if world > 1 and one_graph and not args.no_other_configs and not args.no_allreduce and args.dtype == "fp32":
        other_multi = {}
        n_other = max(3, min(args.steps, 10))

        def max_over_ranks(ms):
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        def all_ranks_ok(ok):
            t = torch.tensor([1.0 if ok else 0.0], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return float(t.item()) > 0.5

        for name, key in (("R50_ovis_720_train_fp32", "r50_720"), ("swinl_ytvis21_fp32", "swinl_360")):
            if key == args.shape:
                continue
            # build and capture first (what can fail), agree across the ranks, only then enter the timed collectives together
            st2 = g2 = g3 = None
            err = None
            try:
                calls2, mask2 = build_calls(torch, SHAPES[key], args.dist, rank, args.layers, device=device)
                st2 = DeviceStep(torch, lib, libmod, calls2, mask2, device, "fp32")
                for _ in range(2):
                    st2.run()
                torch.cuda.synchronize()
                if key == "r50_720":
                    g2 = capture(torch, lambda: whole_step(True, st2))
                    g3 = capture(torch, lambda: whole_step(False, st2))
                else:
                    g2 = capture(torch, lambda: st2.run("fwd"))
            except Exception as e:  # noqa: BLE001 -- a failing side measurement must not lose the headline line
                err = repr(e)[:300]
            if not all_ranks_ok(err is None):
                other_multi[name] = {"error": err or "failed on another rank"}
            elif key == "r50_720":
                sync_all()
                ms2 = max_over_ranks(time_replays(torch, g2, n_other))
                sync_all()
                ms3 = max_over_ranks(time_replays(torch, g3, n_other))
                other_multi[name] = {"workload": workload_text(SHAPES[key]), "n_gpus": world, "train_step_ms": ms2,
                                     "aggregate_train_clips_per_s": world * 1e3 / ms2, "train_step_ms_without_allreduce": ms3,
                                     "what": "one clip per GPU, 36 MSDeformAttn fwd+bwd + mask, gradient all-reduce of the enc+dec parameters "
                                             "(same buckets and kernel as the headline step) inside the one captured graph; max over ranks"}
            else:
                sync_all()
                ms2 = max_over_ranks(time_replays(torch, g2, n_other))
                other_multi[name] = {"workload": workload_text(SHAPES[key]), "n_gpus": world, "inference_ms": ms2,
                                     "aggregate_inference_clips_per_s": world * 1e3 / ms2,
                                     "what": "clips sharded over the GPUs, forward calls of a clip + mask logits per GPU, no collective; max over ranks"}
            g2 = g3 = st2 = calls2 = mask2 = None
            torch.cuda.synchronize()
            torch.cuda.empty_cache()

    # ---- the other BASELINE.json configurations and the module-level arm (single-GPU runs; rank 0 of a multi-GPU run skips them)
    other, modarm = other_multi, modarm_ddp
    if world == 1 and not args.no_other_configs and not args.no_graph:
        del step
        torch.cuda.empty_cache()
        n_other = max(3, min(args.steps, 10))
        other = {}
        for name, key, dt_ in (("R50_ovis_720_train_fp32", "r50_720", "fp32"), ("swinl_ytvis21_fp32", "swinl_360", "fp32"),
                               ("R50_ovis_360_bf16", "r50_360", "bf16"), ("R50_ovis_720_bf16", "r50_720", "bf16")):
            if key == args.shape and dt_ == args.dtype:
                continue
            try:
                other[name] = measure_other_config(torch, lib, libmod, key, dt_, args.dist, device, n_other, hbm_peak)
            except Exception as e:  # noqa: BLE001 -- a failing side measurement must not lose the headline line
                other[name] = {"error": repr(e)[:300]}
        try:
            modarm = module_arm(torch, shape, device, n_other, args.dist)
        except Exception as e:  # noqa: BLE001
            modarm = {"error": repr(e)[:300]}

    # ---- inference window of the reference: the encoder sees a window of up to 30 frames at once
    if other is not None and world == 1 and args.shape == "r50_360":
        try:
            Nw, Sw = 30, sum(h * w for h, w in PYRAMID)
            gw = torch.Generator(device=device).manual_seed(7)
            vw = torch.randn(Nw, Sw, HEADS, HEAD_DIM, device=device, generator=gw)
            refw = ref_points(torch, PYRAMID, device)
            locw = (refw.view(1, Sw, 1, 1, 1, 2) + 0.05 * torch.randn(Nw, Sw, HEADS, len(PYRAMID), POINTS, 2, device=device, generator=gw)).clamp_(-0.1, 1.1)
            aww = torch.softmax(torch.randn(Nw, Sw, HEADS, len(PYRAMID) * POINTS, device=device, generator=gw), -1)
            shw = torch.tensor(PYRAMID, device=device)
            lsw = torch.cat([shw.new_zeros(1), (shw[:, 0] * shw[:, 1]).cumsum(0)[:-1]])
            outw = torch.empty(Nw, Sw, HEADS * HEAD_DIM, device=device)
            run_w = lambda: libmod.check(lib.msda_forward(torch.cuda.current_stream(device).cuda_stream, libmod.MSDA_F32, vw.data_ptr(), shw.data_ptr(), lsw.data_ptr(), locw.data_ptr(),
                                                          aww.data_ptr(), Nw, Sw, HEADS, HEAD_DIM, len(PYRAMID), Sw, POINTS, outw.data_ptr()), "msda_forward")
            w_ms = time_replays(torch, capture(torch, run_w), n_other)
            bytes_w = 4 * (vw.numel() + locw.numel() + aww.numel() + outw.numel())
            other["R50_ovis_360_encoder_window30_fwd"] = {"what": "one encoder MSDeformAttn forward over the reference's 30-frame inference window (N=30, S=Lq=5100)",
                                                         "us": w_ms * 1e3, "frames_per_s": Nw / (w_ms * 1e-3), "algorithmic_bytes": bytes_w,
                                                         "frac_of_hbm_peak": bytes_w / (w_ms * 1e-3) / 1e9 / hbm_peak}
            del vw, locw, aww, outw
        except Exception as e:  # noqa: BLE001
            other["R50_ovis_360_encoder_window30_fwd"] = {"error": repr(e)[:300]}

    # ---- CPU baseline (rank 0, single-GPU runs only): the reference's CPU path restated in torch
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        tl, tm = cpu_sample(torch, calls, mask, reps=3, warm=1)
        cpu = {"value": 1.0 / (N_LAYERS * tl + tm), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": "1 of 6 layers (1 enc + 1 dec-spatial + 4 dec-temporal MSDA fwd+bwd) + mask fwd+bwd via oracle/torch_port.py, "
                         "mean of 3 after 1 warm-up; clip time = 6*layer + mask",
               "layer_ms": tl * 1e3, "mask_ms": tm * 1e3}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic", "config": config_dict(args, shape),
                "launch": ("cuda_graph (one graph per step, NCCL all-reduces captured on a forked branch)" if one_graph else
                           ("cuda_graph" if graphs is not None else "eager")), "launch_note": graph_note,
                "ms_per_step_without_allreduce": no_coll_ms,
                "zero_fill": ("grad_value accumulators zero-filled on a side branch of the step (msda_zero_fill, MSDA_BWD_ACC_ZEROED)" if not args.no_prezero
                              else "inside every backward call"),
                "allreduce": ((("msda_allreduce_f32 (%s over NVLink peer memory, %d CTAs; 16 for the last bucket)" % (peer_ar.algo, args.allreduce_ctas)) if peer_ar is not None else "NCCL")
                              + ", %d buckets per step in gradient-ready order: decoder 14.94M fp32 and one 0.76M bucket per encoder layer, each overlapped with the encoder backward still to run" % (1 + n_enc_layers)) if world > 1 else None,
                "allreduce_note": peer_note,
                # kernels of this library launched inside the timed region: the library's own launch counter over an eager pass of
                # the same step (the graph replays re-launch exactly those kernels) x steps; `launches_per_step` is the C-ABI call count
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(round(counted_per_step * args.steps)) if counted_per_step else gpu_launches,
                "launches_per_step": launches_per_step,
                "lib_launch_count_per_eager_step": counted_per_step,
                "roofline": roofline, "roofline_binding_resource": binding, "roofline_l2_reductions": l2_reduction,
                "roofline_fwd": roofline_fwd, "roofline_fwd_binding_resource": binding_fwd, "roofline_mask": roofline_mask,
                "roofline_mask_bwd": roofline_mask_bwd, "eager_ms_per_step": eager_ms, "step_with_inline_zero_fill": inline_fill,
                "forward_only": fwd_only, "other_configs": other, "module_arm": modarm, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1 and peer_ar is not None:
        torch.cuda.synchronize()
        peer_ar.check()
    if world > 1:
        # captured NCCL kernels keep references into the communicator: drop the graphs and drain the device before tearing the
        # process group down, and do not let a slow teardown hold the line (already printed) hostage
        graphs = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        import threading
        t = threading.Thread(target=dist.destroy_process_group, daemon=True)
        t.start()
        t.join(20.0)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
