#!/usr/bin/env python
"""bench.py — BASELINE.json metric on synthetic data: RTFx (audio-seconds per second).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA sm_100a through the C ABI)
  python bench.py --impl reference --gpus N ...            the reference path's CPU restatement (oracle port) on the host cores

Workloads:
  --workload transcribe (default)  BASELINE.json configs[2]: large-v3 (128 mel bins) full transcribe + DTW of 1 h of synthetic
                                   audio = 120 x 30 s windows PER GPU, sharded by window (no data-path collective, weak scaling):
                                   log-mel -> encoder -> cross-KV -> greedy decode -> token timestamps -> DTW, through wdr_full_batch_*.
  --workload encoder               BASELINE.json configs[1]: tiny.en batched log-mel + encoder over 64 x 30 s windows on one GPU.
  --workload diarize               BASELINE.json configs[3]: segmentation-3.0 windows + WeSpeaker ResNet34 embeddings + cosine matrix +
                                   clustering on a 10 min 4-speaker synthetic mix per GPU (one recording per rank).

One JSON line on stdout (rank 0).  `value` = whole-job RTFx with the PCM already resident in HBM (wdr_full_batch_i16_dev); `e2e` =
the same through the host-pointer C ABI call (pinned host PCM -> H2D -> ... -> results read back through the whisper.h-style
accessors); `roofline` = the dominant kernel class, timed with CUDA events on the launching stream inside the timed region;
`cpu_baseline` = the oracle port on a bounded sample of the same workload.
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

ARCH_DIMS = {"tiny.en": (384, 6, 4, 4, 80), "base.en": (512, 8, 6, 6, 80), "small": (768, 12, 12, 12, 80),
             "large-v3": (1280, 20, 32, 32, 128), "large-v3-turbo": (1280, 20, 32, 4, 128)}


def flops_per_window(arch):
    """Algorithmic FLOPs (2*M*N*K; softmax/LN/GELU excluded) of one 30 s window, split by kernel class."""
    d, _, L, Ld, n_mel = ARCH_DIMS[arch]
    T = 1500
    conv = 2 * 3000 * d * 3 * n_mel + 2 * T * d * 3 * d
    lin = L * 24 * T * d * d
    att = L * 4 * T * T * d
    cross = Ld * 2 * T * d * 2 * d
    return {"gemm_enc": conv + lin, "attention": att, "cross_kv": cross, "total_enc": conv + lin + att}


def synth_pcm(n_chunks, seed0=2000):
    """Seeded synthetic 16 kHz int16 'speech' (tests/conftest.py:synth_audio), one 30 s window per seed."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import synth_audio
    base = [synth_audio(seed0 + i, 30.0) for i in range(min(n_chunks, 4))]
    out = np.empty((n_chunks, 480000), np.int16)
    for i in range(n_chunks):  # 4 distinct windows, rolled so that no two windows are identical
        out[i] = np.roll(base[i % len(base)], 7919 * (i // len(base)))
    return out


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU port (the oracle) — the reported baseline and the --impl reference arm
# ---------------------------------------------------------------------------------------------------------------------
def host_cores():
    """Cores this process may run on (the affinity mask, not the machine's count)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def claim_cpu_threads(n=None):
    """The CPU arm sets its thread count ITSELF: a launcher may export OMP_NUM_THREADS=1 (torch.distributed.run does for
    multi-rank launches), which would run the port on one thread while the line claims all cores.  Returns the team size an OpenMP
    parallel region of the port really gets — the number reported as `cores`."""
    from oracle import native
    n = n or host_cores()
    got = native.set_threads(n)
    try:  # the numpy / torch parts of the diarization port (oracle/pyannet.py, oracle/resnet.py)
        import torch
        torch.set_num_threads(n)
    except Exception:
        pass
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
    except Exception:
        pass
    return got


class CpuPort:
    def __init__(self, arch, workload):
        from oracle import native, weights as W, filters
        native.build()
        self.threads = claim_cpu_threads()
        self.native, self.W = native, W
        self.arch, self.workload = arch, workload
        self.a = W.ARCHS[arch]
        w = W.whisper_weights(arch, seed=1234)
        self.enc_w = W.pack_encoder(arch, w)
        self.filt = filters.whisper_mel_filters(self.a["n_mel"])
        self.dec = native.Decoder(arch, W.pack_decoder(arch, w), bf16=True) if workload == "transcribe" else None

    def run(self, pcm, n_chunks):
        """Processes n_chunks windows; returns (RTFx, seconds, tokens).  The result of window 0 is kept in self.last0."""
        from oracle import full
        t0 = time.perf_counter()
        n_tok = 0
        for i in range(n_chunks):
            x = pcm[i % len(pcm)].astype(np.float32) / np.float32(32768.0)
            mel = self.native.log_mel(x, self.filt)[:, :3000]
            enc = self.native.whisper_encode(np.ascontiguousarray(mel), self.arch, self.enc_w)
            if self.dec is not None:
                r = full.full_window(self.dec, enc, x)
                n_tok += r.get("n_sampled", 0)
                if i == 0:
                    self.last0 = r
        dt = time.perf_counter() - t0
        return n_chunks * 30.0 / dt, dt, n_tok

    def window_on_encoder_output(self, pcm_row, enc_out, beam_size=1):
        """The decode half (greedy / beam decode, token timestamps, DTW) of one window on a GIVEN encoder output: the strict checker
        of what the GPU arm produced for that window (identical inputs to the decoder => ids / t0 / t1 / t_dtw must be identical)."""
        from oracle import full
        x = pcm_row.astype(np.float32) / np.float32(32768.0)
        return full.full_window(self.dec, enc_out, x, beam_size=beam_size)


# ---------------------------------------------------------------------------------------------------------------------
# diarization workload (BASELINE configs[3])
# ---------------------------------------------------------------------------------------------------------------------
def diar_cpu(pcm, n_windows, n_segments):
    """Oracle port of the diarization path on a bounded sample: n_windows segmentation windows + n_segments embeddings (of the
    segments the library found) + clustering.  Returns (seconds_seg_per_window, seconds_emb_per_frame)."""
    from oracle import native, pyannet, resnet
    native.build()
    wseg = pyannet.pyannet_weights(1234)
    wres = resnet.resnet_weights(1234)
    t0 = time.perf_counter()
    for i in range(n_windows):
        pyannet.pyannet_forward(pcm[i * 160000:(i + 1) * 160000].astype(np.float32), wseg)
    t_seg = (time.perf_counter() - t0) / max(n_windows, 1)
    frames = 0
    t0 = time.perf_counter()
    for i in range(n_segments):
        seg = pcm[i * 48000:(i + 1) * 48000]
        resnet.compute(seg, wres, native.kaldi_fbank)
        frames += 1 + (len(seg) - 400) // 160
    t_emb = (time.perf_counter() - t0) / max(frames, 1)
    return t_seg, t_emb


def run_diarize(args):
    """One step = one 10 min recording: pyannote_rs::get_segments -> EmbeddingExtractor::compute per segment -> cosine matrix ->
    leader scan (the crate's EmbeddingManager policy) + agglomerative labels.  value: PCM already in HBM, stage calls through the
    device-pointer C ABI; e2e: host PCM through whisper-diarize-rs_b200.host.diarize (the crate-shaped call)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import synth_audio
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    seconds = 600.0
    pcm = synth_audio(4001 + rank, seconds, n_speakers=4)
    if args.impl == "reference":
        if rank != 0:
            return
        cores = claim_cpu_threads()
        n_w, n_s = 2, 4
        t0 = time.perf_counter()
        steps = max(1, min(args.steps, 3))
        for _ in range(steps):
            t_seg, t_emb = diar_cpu(pcm, n_w, n_s)
        dt = (time.perf_counter() - t0) / steps
        audio_s = n_w * 10.0
        val = audio_s / dt
        line = {"impl": "reference", "metric": "RTFx diarization", "value": val, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": steps, "warmup": 0,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "diarization of a 10 min 4-speaker synthetic mix (BASELINE configs[3])"},
                "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port",
                                 "sample": f"{n_w} segmentation windows (20 s of audio) + {n_s} x 3 s segment embeddings per step, oracle/pyannet.py + oracle/resnet.py"},
                "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    import wdr_b200 as w
    from hostmirror import host as H
    if w.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: libwdr_b200 has no CPU path")
    torch.cuda.set_device(local)
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    seg = w.Segmenter(seed=1234, device=local)
    emb = w.EmbeddingExtractor(seed=1234, device=local)
    pcm_pin = torch.from_numpy(pcm).pin_memory()
    out = {}
    wd = w.dist.connect(w, local) if world > 1 else None  # the exchange itself is the library's (csrc/dist.cu, NCCL); torch only carries the id
    gathered = torch.empty(world * 4096, emb.dim, device="cuda", dtype=torch.float32) if world > 1 else None
    mgr_box = {}

    def step():
        if world == 1:
            out["r"] = H.diarize(seg, emb, pcm_pin.numpy(), 0.5, w.SIZE_MAX, "leader")
            return
        # N > 1: every rank holds one 10 min shard of a (world x 10 min) recording; segments and embeddings are local, the
        # embeddings are all-gathered (NCCL over NVLink: the path's one exchange step) and clustered globally on every rank
        segs_l = seg.get_segments(pcm_pin.numpy())
        off_l = np.zeros(len(segs_l) + 1, np.int64)
        for i, sg_ in enumerate(segs_l):
            off_l[i + 1] = off_l[i] + len(sg_["samples"])
        cat_l = np.concatenate([sg_["samples"] for sg_ in segs_l]).astype(np.int16) if len(segs_l) else np.zeros(0, np.int16)
        # embeddings stay on the device: segment PCM H2D -> fbank + ResNet34 -> rows of E_dev; the rows of embeddable segments are
        # L2-normalised into the send buffer and all-gathered by the library (wdr_allgather_embeddings_dev); one D2H of the global
        # table; the crate's speaker policy (wdr_spk_assign_batch) runs on it on every rank: O(n x speakers x D), no n x n matrix
        pcm_d = torch.from_numpy(cat_l).pin_memory().cuda(non_blocking=True)
        E_dev = torch.empty(max(len(segs_l), 1), emb.dim, device="cuda", dtype=torch.float32)
        st_l = emb.compute_batch_dev(pcm_d.data_ptr(), off_l, E_dev.data_ptr(), torch.cuda.current_stream().cuda_stream) if len(segs_l) else np.zeros(0, np.int32)
        ok_l = np.flatnonzero(st_l == 0)
        E_ok = E_dev[torch.from_numpy(ok_l).cuda()].contiguous() if len(ok_l) else E_dev[:0]
        n_all, counts = wd.allgather_embeddings_dev(E_ok.data_ptr(), int(len(ok_l)), emb.dim, gathered.data_ptr(), gathered.shape[0], normalize=True,
                                                    stream=torch.cuda.current_stream().cuda_stream)
        E_all = gathered[:n_all].cpu().numpy()
        mgr = w.EmbeddingManager(w.SIZE_MAX)
        out["r"] = mgr.assign_batch(E_all, 0.5)
        mgr.close()
        out["n_global"] = int(n_all)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = w.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    barrier()
    dt = time.perf_counter() - t0
    launches = w.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    # stage timings of one more step (host wall clock around each blocking C-ABI call)
    stages = {}
    t = time.perf_counter(); segs = seg.get_segments(pcm); stages["segmentation_ms"] = (time.perf_counter() - t) * 1e3
    off = np.zeros(len(segs) + 1, np.int64)
    for i, sgm in enumerate(segs):
        off[i + 1] = off[i] + len(sgm["samples"])
    cat = np.concatenate([sgm["samples"] for sgm in segs]).astype(np.int16) if len(segs) else np.zeros(0, np.int16)
    t = time.perf_counter(); E, status = emb.compute_batch(cat, off); stages["embedding_ms"] = (time.perf_counter() - t) * 1e3
    flops = emb.last_flops()
    emb.profile(True)   # one more call with a CUDA-event pair around every GEMM / im2col gather: the GEMM-only roofline figure
    emb.compute_batch(cat, off)
    gemm_ms, gather_ms = emb.last_kernel_ms()
    emb.profile(False)
    stages["embedding_gemm_ms"], stages["embedding_im2col_ms"] = gemm_ms, gather_ms
    ok = np.flatnonzero(status == 0)
    if not len(ok):
        raise SystemExit("diarize workload: the segmenter emitted no embeddable segment")
    t = time.perf_counter(); S = w.cosine_matrix(E[ok]); lab = w.cluster_leader(S, 0.5); agg = w.cluster_agglomerative(S, 0.5)
    stages["cosine_cluster_ms"] = (time.perf_counter() - t) * 1e3
    # Kaldi fbank kernel alone on the recording's segments, device pointers, CUDA events on the launching stream (north-star: the
    # mel / fbank kernels are quoted in HBM GB/s).  Algorithmic bytes: int16 samples in + 80 fp32 bins per frame out.
    fb = {}
    try:
        live = [i for i in ok]
        so = np.zeros(len(live) + 1, np.int64)
        fo = np.zeros(len(live) + 1, np.int64)
        for k, i in enumerate(live):
            n_i = int(off[i + 1] - off[i])
            so[k + 1] = so[k] + n_i
            fo[k + 1] = fo[k] + (1 + (n_i - 400) // 160)
        cat_live = np.concatenate([segs[i]["samples"] for i in live]).astype(np.int16)
        d_pcm = torch.from_numpy(cat_live).cuda()
        d_so, d_fo = torch.from_numpy(so).cuda(), torch.from_numpy(fo).cuda()
        d_out = torch.empty(int(fo[-1]) * 80, device="cuda", dtype=torch.float32)
        stream = torch.cuda.current_stream().cuda_stream
        L = w.load()
        call = lambda: L.wdr_kaldi_fbank_batch_i16_dev(d_pcm.data_ptr(), d_so.data_ptr(), d_fo.data_ptr(), len(live), int(fo[-1]), 80, 1, d_out.data_ptr(), stream)
        for _ in range(3):
            assert call() == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms_fb = e0.elapsed_time(e1) / 10
        nbytes = cat_live.nbytes + int(fo[-1]) * 80 * 4
        fb = {"ms": ms_fb, "frames": int(fo[-1]), "algorithmic_bytes": int(nbytes), "achieved_gbs": nbytes / (ms_fb / 1e3) / 1e9,
              "note": "fbank + per-segment mean subtraction incl. the call's offset read-back; fp32 512-point FFT per frame: ALU-bound like the log-mel kernel"}
    except Exception as ex:  # the bench line must not die on the auxiliary measurement
        fb = {"error": str(ex)}
    if world > 1:
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    value = world * seconds * args.steps / dt
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        ach = flops / (stages["embedding_ms"] / 1e3) / 1e12 if stages["embedding_ms"] > 0 else 0.0
        ach_gemm = flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        line = {"metric": "RTFx diarization", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16/f32", "data": "synthetic",
                "config": {"workload": "diarization of a 10 min 4-speaker synthetic mix per GPU: segmentation-3.0 windows (fp32) + WeSpeaker ResNet34 "
                                       "embeddings (bf16 tcgen05 GEMMs) + cosine matrix + leader scan (BASELINE configs[3])",
                           "windows": int(w.load().wdr_seg_n_windows(len(pcm))), "segments": len(segs), "embedded": int(len(ok)),
                           "speakers_leader": int(lab.max()) if len(lab) else 0, "clusters_agglomerative": int(agg.max()) if len(agg) else 0,
                           "l2": "each step streams the whole recording's activations (> 126 MB L2)",
                           "exchange": None if world == 1 else f"all-gather of {out.get('n_global')} x {emb.dim} fp32 L2-normalised embeddings per step inside the library "
                                                                f"(wdr_allgather_embeddings_dev, NCCL {w.load().wdr_dist_nccl_version()}), the crate's speaker policy on the global table on every rank"},
                "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": int(pcm.nbytes + cat.nbytes), "d2h_bytes_per_step": int(E.nbytes + 60 * 589 * 7 * 4),
                        "api": "host.diarize: wdr_seg_get_segments + wdr_emb_compute_batch_i16 + wdr_cosine_matrix + wdr_cluster_leader (host pointers)"},
                "gpu_launches": int(launches), "clocks": clocks, "stages": stages, "fbank_kernel": fb,
                "roofline": {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05): the 36 convolutions of the ResNet34 embedding stage — 29 as implicit GEMMs over the zero-padded "
                                                          "activation maps, the stem and the stride-2 ones over a gathered operand (CUDA events around the GEMM launches only)",
                             "achieved": ach_gemm, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach_gemm / tf_peak, "traffic": None,
                             "algorithmic_gflop_per_step": flops / 1e9, "gemm_ms": gemm_ms, "im2col_gather_ms": gather_ms,
                             "whole_stage": {"achieved": ach, "frac": ach / tf_peak, "note": "whole embedding call incl. fbank, operand gathers of the stem / stride-2 stages, H2D / D2H"}}}
        if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N = 1 only
            cpu_threads = claim_cpu_threads()
            t_seg, t_emb = diar_cpu(pcm, 2, 4)
            frames = sum(1 + (int(off[i + 1] - off[i]) - 400) // 160 for i in ok)
            est = 60 * t_seg + frames * t_emb
            line["cpu_baseline"] = {"value": seconds / est, "unit": "audio-s/s", "cores": cpu_threads, "kind": "port",
                                    "sample": f"2 of 60 segmentation windows ({t_seg:.2f} s each) + 4 x 3 s embeddings ({t_emb * 1e3:.2f} ms per fbank frame), "
                                              f"extrapolated to the recording's 60 windows and {frames} frames; oracle/pyannet.py (numpy) + oracle/resnet.py (torch CPU)"}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    seg.close()
    emb.close()
    if wd is not None:
        wd.close()
    if world > 1:
        dist.destroy_process_group()


def run_pipeline(args):
    """BASELINE configs[4] in miniature: the whole crate-shaped flow on one recording per GPU — Silero VAD mask, pyannote
    segmentation -> SpeechSegments, every segment through state.full (large-v3-turbo, token + DTW timestamps) and the WeSpeaker
    embedding, EmbeddingManager speaker ids, subtitle cues — through whisper-diarize-rs_b200.host with HOST buffers (there is no
    device-resident arm: the step IS the public API call, so value == e2e).  The full config is 8 h over 8 GPUs (1 h per GPU);
    the default here is a 10 min recording per GPU so that the run finishes in minutes (--minutes)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import synth_audio
    import torch
    import torch.distributed as dist
    import wdr_b200 as w
    from hostmirror import host as H
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if w.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: libwdr_b200 has no CPU path")
    torch.cuda.set_device(local)
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    seconds = 60.0 * args.minutes
    pcm = synth_audio(5001 + rank, seconds, n_speakers=4)
    arch = args.arch
    ctx = w.Context(arch, seed=1234, gpu_device=local, enable_dtw=True)
    st = ctx.create_state()
    vad = w.VadContext(seed=1234, gpu_device=local)
    seg = w.Segmenter(seed=1234, device=local)
    emb = w.EmbeddingExtractor(seed=1234, device=local)
    out = {}

    def step():
        t = time.perf_counter()
        mask, _ = H.vad_get_segments(vad, pcm)
        t1 = time.perf_counter()
        speech = [dict(start=s_["start"], end=s_["end"], samples=s_["samples"]) for s_ in seg.get_segments(pcm)]
        t2 = time.perf_counter()
        segs, lang = H.run_transcription_pipeline_sharded(st, speech, None, emb, 0.5, w.SIZE_MAX)
        t3 = time.perf_counter()
        cues = H.format_cues(segs, lang, mask)
        t4 = time.perf_counter()
        out.update(mask=len(mask), speech=len(speech), segments=len(segs), cues=len(cues), speakers=len(set(s_["speaker_id"] for s_ in segs)),
                   stages={"vad_ms": (t1 - t) * 1e3, "segmentation_ms": (t2 - t1) * 1e3, "transcribe_embed_ms": (t3 - t2) * 1e3, "format_ms": (t4 - t3) * 1e3},
                   pcm_bytes=int(pcm.nbytes + sum(len(s_["samples"]) for s_ in speech) * 2 * 2))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = w.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    barrier()
    dt = time.perf_counter() - t0
    launches = w.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    value = world * seconds * args.steps / dt
    # one more step with a CUDA-event pair around every launch of the transcription state (wdr_profile_*): the kernel-class times
    # behind the roofline block (the encoder dominates large-v3-turbo: 32 encoder layers against 4 decoder layers)
    st.profile_enable(True)
    st.profile_collect()
    step()
    torch.cuda.synchronize()
    prof = st.profile_collect()
    st.profile_enable(False)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        fl = flops_per_window(arch)
        n_win = out["speech"]  # every speech segment is its own 30 s whisper window
        roofline, kern = None, {}
        if prof["gemm"]["ms"] > 0:
            ach = (fl["gemm_enc"] + fl["cross_kv"]) * n_win / (prof["gemm"]["ms"] / 1e3) / 1e12
            roofline = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05): encoder + cross-KV GEMMs of the transcription stage", "achieved": ach, "peak": tf_peak,
                        "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": None, "windows": n_win,
                        "share_of_profiled_kernel_time": prof["gemm"]["ms"] / max(sum(v["ms"] for v in prof.values()), 1e-9),
                        "peak_source": "measured (MEASURED_PEAKS.json: sustained bf16)" if peaks else "fallback (B200_PROFILING.md)"}
        for name, v in prof.items():
            if v["ms"] > 0:
                kern[name] = {"ms_per_step": v["ms"], "launches_per_step": v["records"]}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            # the CPU port on a bounded sample of the same recording: VAD on 10 s, 1 segmentation window, 2 speech segments through
            # mel + encoder + decode + DTW (large-v3-turbo) and the ResNet34 embedding; extrapolated per stage to the recording
            from oracle import vad as OV
            threads = claim_cpu_threads()
            x10 = pcm[:160000].astype(np.float32) / np.float32(32768.0)
            t = time.perf_counter(); OV.silero_probs(x10, OV.vad_weights(1234)); t_vad = (time.perf_counter() - t) / 10.0
            t_seg, t_emb = diar_cpu(pcm, 1, 2)
            port = CpuPort(arch, "transcribe")
            win = np.zeros((2, 480000), np.int16)
            take = [s_ for s_ in seg.get_segments(pcm)][:2]
            for i_, s_ in enumerate(take):
                win[i_, : min(len(s_["samples"]), 480000)] = s_["samples"][:480000]
            _, t_asr, _ = port.run(win, max(1, len(take)))
            t_asr /= max(1, len(take))
            frames = sum(max(0, 1 + (len(s_["samples"]) - 400) // 160) for s_ in seg.get_segments(pcm))
            est = seconds * t_vad + (seconds / 10.0) * t_seg + n_win * t_asr + frames * t_emb
            cpu = {"value": seconds / est, "unit": "audio-s/s", "cores": threads, "kind": "port",
                   "sample": f"VAD on 10 s ({t_vad * 1e3:.1f} ms per audio-second), 1 segmentation window ({t_seg:.2f} s), {len(take)} speech segments through "
                             f"mel + encoder + decode + DTW ({t_asr:.1f} s each) and 2 x 3 s embeddings ({t_emb * 1e3:.2f} ms per fbank frame), extrapolated to "
                             f"{seconds:.0f} s of audio, {n_win} speech segments, {frames} fbank frames; oracle/*.py + oracle/wdr_oracle*.c"}
        line = {"metric": "RTFx VAD + diarization + transcribe", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16/f32", "data": "synthetic",
                "config": {"workload": f"{arch} transcribe + diarization + Silero VAD of a {args.minutes} min 4-speaker synthetic recording per GPU through the "
                                       "crate-shaped host flow, speech segments batched 128 per wdr_full_batch_i16 call (BASELINE configs[4], bounded)",
                           "speech_segments": out["speech"], "vad_ranges": out["mask"], "segments": out["segments"], "cues": out["cues"], "speakers": out["speakers"],
                           "note": "every speech segment is its own 30 s whisper window, as in the reference; the random-init segmentation net emits "
                                   "many short segments, so windows per audio-hour are far above a real recording's",
                           "l2": "each step streams the whole recording's activations (> 126 MB L2)"},
                "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": out["pcm_bytes"], "d2h_bytes_per_step": None,
                        "api": "host.vad_get_segments + Segmenter.get_segments + host.run_transcription_pipeline_sharded + host.format_cues"},
                "gpu_launches": int(launches), "clocks": clocks, "stages_last_step": out["stages"], "roofline": roofline, "kernels": kern, "cpu_baseline": cpu}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    for m in (vad, seg, emb, st, ctx):
        m.close()
    if world > 1:
        dist.destroy_process_group()


def parity_check(w, st, ctx, port, pcm_host, params, beam_size=1):
    """bench.py checks what it times: window 0 of the timed batch against the CPU port.
    strict: the port's decode half (greedy decode, A.5 token timestamps, DTW) on the LIBRARY's encoder output for that window —
            identical decoder inputs, so token ids, segment times, t0 / t1 and t_dtw must be identical;
    from_pcm: the port end to end in fp32 from the PCM (its own mel and fp32 encoder, already computed by the cpu_baseline sample) —
            a divergence there is legitimate only where the fp32 model's own top-1 margin is below the bf16 noise floor."""
    L = w.load()
    hid0 = st.encode_chunks(pcm_host[0:1])[0]
    st.full_batch(pcm_host[0:1], None, params)  # window 0 alone: results are independent of the batch composition (tests/test_gpu_full_size.py)
    got = [s_ for s_ in st.segments() if s_["chunk"] == 0]
    ids_g = [t.id for s_ in got for t in s_["tokens"]]
    out = {"window": 0, "n_tokens": len(ids_g)}
    ref = port.window_on_encoder_output(pcm_host[0], hid0, beam_size=beam_size)
    toks_r = [t for s_ in ref["segments"] for t in s_["tokens"]]
    ids_r = [t.id for t in toks_r]
    same = ids_g == ids_r
    out["tokens_identical"] = bool(same)
    if same:
        toks_g = [t for s_ in got for t in s_["tokens"]]
        out["first_divergence"] = None
        out["segment_times_identical"] = [(s_["t0"], s_["t1"]) for s_ in got] == [(s_["t0"], s_["t1"]) for s_ in ref["segments"]]
        out["t0_t1_identical"] = [(t.t0, t.t1) for t in toks_g] == [(t.t0, t.t1) for t in toks_r]
        out["t_dtw_identical"] = [t.t_dtw for t in toks_g] == [t.t_dtw for t in toks_r]
    else:
        k = next((i for i, (a_, b_) in enumerate(zip(ids_g, ids_r)) if a_ != b_), min(len(ids_g), len(ids_r)))
        out["first_divergence"] = int(k)
        out["margin"] = float(ref["margins"][k]) if k < len(ref["margins"]) else None
    fp = getattr(port, "last0", None) if beam_size <= 1 else None  # the cpu_baseline sample is a greedy run
    if fp is not None:
        ids_f = [t.id for s_ in fp["segments"] for t in s_["tokens"]]
        out["from_pcm_fp32_encoder"] = {"tokens_identical": ids_f == ids_g}
        if ids_f != ids_g:
            k = next((i for i, (a_, b_) in enumerate(zip(ids_g, ids_f)) if a_ != b_), min(len(ids_g), len(ids_f)))
            out["from_pcm_fp32_encoder"].update(first_divergence=int(k), oracle_top1_margin=float(fp["margins"][k]) if k < len(fp["margins"]) else None)
    out["checker"] = ("oracle/full.py:full_window on oracle/wdr_oracle_full.c (library storage mode: bf16 cross-KV, f16 self-KV)" +
                      (f", beam search (beam {beam_size})" if beam_size > 1 else ", greedy") + ", window 0 of the timed batch")
    return out


def metric_name(workload):
    return "RTFx full transcribe+DTW" if workload == "transcribe" else "RTFx mel+encoder"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pcm = synth_pcm(4)
    port = CpuPort(args.arch, args.workload)
    cores = port.threads  # the OpenMP team size the port really runs with (set explicitly; OMP_NUM_THREADS of a launcher is overridden)
    _, t1, _ = port.run(pcm, 1)  # warm (page-in) and calibrate the per-step sample
    per_step = max(1, min(4, int(8.0 / max(t1, 1e-3))))
    for _ in range(max(args.warmup - 1, 0)):
        if t1 < 20.0:
            port.run(pcm, 1)
    steps = args.steps if t1 * per_step * args.steps < 240.0 else max(1, int(240.0 / (t1 * per_step)))
    t0 = time.perf_counter()
    for _ in range(steps):
        port.run(pcm, per_step)
    dt = time.perf_counter() - t0
    val = steps * per_step * 30.0 / dt
    line = {"impl": "reference", "metric": metric_name(args.workload), "value": val, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "chunks_per_step": per_step},
            "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port",
                             "sample": f"{per_step} x 30 s window(s) per step x {steps} steps on {cores} OpenMP threads (omp_get_num_threads() inside a parallel region; "
                                       f"{host_cores()} cores in the affinity mask), oracle/wdr_oracle*.c"},
            "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(args):
    if args.workload == "transcribe":
        return (f"{args.arch} full transcribe (log-mel + encoder + cross-KV + greedy decode + token timestamps + DTW) of "
                f"{args.chunks} x 30 s windows per GPU, sharded by window (BASELINE configs[2])")
    return f"{args.arch} batched log-mel + encoder over {args.chunks} x 30 s windows per GPU (BASELINE configs[1])"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="transcribe", choices=["transcribe", "encoder", "diarize", "pipeline"])
    ap.add_argument("--minutes", type=int, default=10, help="pipeline workload: recording length per GPU")
    ap.add_argument("--arch", default=None)
    ap.add_argument("--chunks", type=int, default=None, help="30 s windows per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--beam", type=int, default=0, help="transcribe workload: beam search with this beam size (the crate's default strategy, beam 5) instead of greedy")
    ap.add_argument("--temperature-inc", type=float, default=0.0, help="transcribe workload: whisper_full's temperature ladder increment (whisper.cpp default 0.2; 0 = no fallback)")
    args = ap.parse_args()
    if args.arch is None:
        args.arch = "large-v3" if args.workload == "transcribe" else "large-v3-turbo" if args.workload == "pipeline" else "tiny.en"
    if args.chunks is None:
        args.chunks = 120 if args.workload == "transcribe" else 64
    if args.steps is None:
        args.steps = 5 if args.workload == "transcribe" else 40 if args.workload == "encoder" else 10
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.workload == "diarize":
        if args.steps is None or args.steps > 40:
            args.steps = 10
        return run_diarize(args)
    if args.workload == "pipeline":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the pipeline workload has no CPU arm; see the transcribe and diarize workloads"}))
            return
        if args.steps is None or args.steps > 10:
            args.steps = 3
        return run_pipeline(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import wdr_b200 as w

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:  # one process per GPU shares the box's cores: split them instead of 16 host threads per rank
        os.environ.setdefault("WDR_HOST_THREADS", str(max(2, (os.cpu_count() or 16) // world)))
    if w.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: libwdr_b200 has no CPU path")
    torch.cuda.set_device(local)
    # stdout carries exactly ONE JSON line: anything libraries print while we run (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()

    B = args.chunks
    full = args.workload == "transcribe"
    ctx = w.Context(args.arch, seed=1234, gpu_device=local, enable_dtw=full)
    st = ctx.create_state()
    d = ctx.dims.n_audio_state
    pcm_host = synth_pcm(B, seed0=2000 + 64 * rank)
    pcm_pin = torch.from_numpy(pcm_host).pin_memory()
    pcm_dev = pcm_pin.cuda(non_blocking=False)
    stream = torch.cuda.current_stream().cuda_stream
    params = None
    if full:
        kw = dict(temperature_inc=args.temperature_inc)
        params = st.full_params(strategy=1, beam_size=args.beam, **kw) if args.beam > 1 else st.full_params(**kw)
    hidden = None if full else torch.empty(B, 1500, d, device="cuda", dtype=torch.float32)
    stats = {"segments": 0, "tokens": 0}

    def step_resident():
        if full:
            stats["segments"] = st.full_batch_dev(pcm_dev.data_ptr(), B, 480000, params)
        else:
            st.encode_chunks_dev(pcm_dev.data_ptr(), 480000, B, hidden.data_ptr(), None, stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ----
    # Timed region 1 (-> value): K steps exactly as a caller runs them (decode iterations replay as CUDA graphs, no per-kernel events).
    # Timed region 2 (-> roofline / kernels): the same K steps again with a CUDA-event pair around every launch on the launching
    # stream (wdr_profile_*); graphs are off there because events cannot bracket nodes of a replayed graph.
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = w.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()  # no-op unless run under `ncu --profile-from-start off`
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = w.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    phases = st.phase_ms() if full else None  # of the last timed step
    st.profile_enable(True)
    st.profile_collect()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        step_resident()
    p1.record()
    barrier()
    ms_profiled = p0.elapsed_time(p1)
    prof = st.profile_collect()
    st.profile_enable(False)
    cross_launches, cross_live = st.cross_attn_stats() if full else (0, 0)  # of the last step, counted on the device
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * 30.0 * args.steps / (ms / 1e3)

    # ---- strong scaling (N > 1): BASELINE configs[2] read literally — args.chunks windows in TOTAL, chunks / N per GPU ----
    strong = None
    if full and world > 1:
        Bs = max(1, B // world)
        for _ in range(2):
            st.full_batch_dev(pcm_dev.data_ptr(), Bs, 480000, params)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            st.full_batch_dev(pcm_dev.data_ptr(), Bs, 480000, params)
        s1.record()
        barrier()
        ms_s = s0.elapsed_time(s1)
        t = torch.tensor([ms_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_s = float(t.item())
        strong = {"scaling": "strong", "windows_total": Bs * world, "windows_per_gpu": Bs, "value": world * Bs * 30.0 * args.steps / (ms_s / 1e3),
                  "unit": "audio-s/s", "ms_per_step": ms_s / args.steps, "phases_ms_last_step": st.phase_ms(),
                  "note": "same K steps, device-resident PCM, CUDA events, max over ranks; divide by the N=1 weak value for the strong-scaling efficiency"}

    # ---- end-to-end arm: host PCM through the C ABI, H2D inside the timed region, results read back ----
    def step_e2e():
        if full:
            n = st.full_batch_ptr(pcm_pin.data_ptr(), B, 480000, params)
            L = w.load()
            tok = 0
            for i in range(n):  # whisper.h-style accessors, as the crate walks them (src/transcribe.rs:397-412, 252-282)
                L.wdr_full_get_segment_t0_from_state(st._h, i)
                L.wdr_full_get_segment_text_from_state(st._h, i)
                tok += L.wdr_full_n_tokens_from_state(st._h, i)
            stats["tokens"] = tok
            return n
        st.encode_chunks_resident(pcm_pin.data_ptr(), B)
        return st.hidden_digest(B)

    for _ in range(2):
        res = step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = world * B * 30.0 * args.steps / e2e_s
    if full:
        check = {"segments_e2e": int(res), "segments_resident": int(stats["segments"]), "tokens": int(stats["tokens"]),
                 "same_segment_count": bool(res == stats["segments"])}
        d2h = int(B * 224 * 56 + B * 48 + B * 224 * 16)  # token data + decoder state + the A.5 token times (the energy envelope stays in HBM)
    else:
        ref_dig = hidden.abs().mean(dim=(1, 2)).cpu().numpy()
        check = {"digest_matches_resident_arm": bool(np.allclose(res, ref_dig, rtol=1e-3))}
        d2h = 4 * B

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json: sustained bf16, copy HBM)" if peaks else "fallback (B200_PROFILING.md)"
        fl = flops_per_window(args.arch)
        dm, H, L_enc, L_dec, n_mel = ARCH_DIMS[args.arch]
        total_ms = max(sum(v["ms"] for v in prof.values()), 1e-9)
        kern = {}
        # encoder-side GEMMs (conv stem + linears; plus the cross-KV projection in the transcribe workload)
        gemm_flops = fl["gemm_enc"] + (fl["cross_kv"] if full else 0)
        if prof["gemm"]["ms"] > 0:
            ach = gemm_flops * B * args.steps / (prof["gemm"]["ms"] / 1e3) / 1e12
            kern["gemm"] = {"ms_per_step": prof["gemm"]["ms"] / args.steps, "launches_per_step": prof["gemm"]["records"] / args.steps,
                            "achieved_tflops": ach, "frac": ach / tf_peak, "algorithmic_gflop_per_window": gemm_flops / 1e9}
        if prof["attention"]["ms"] > 0:
            ach = fl["attention"] * B * args.steps / (prof["attention"]["ms"] / 1e3) / 1e12
            kern["attention"] = {"ms_per_step": prof["attention"]["ms"] / args.steps, "launches_per_step": prof["attention"]["records"] / args.steps,
                                 "achieved_tflops": ach, "frac": ach / tf_peak}
        mel_bytes = B * (480000 * 2 + n_mel * 3000 * 4)
        if prof["mel"]["ms"] > 0:
            gbs = mel_bytes * args.steps / (prof["mel"]["ms"] / 1e3) / 1e9
            kern["mel"] = {"ms_per_step": prof["mel"]["ms"] / args.steps, "achieved_gbs": gbs, "frac_hbm": gbs / hbm_peak,
                           "algorithmic_bytes_per_window": 480000 * 2 + n_mel * 3000 * 4}
        if prof["dec_cross"]["ms"] > 0:
            # algorithmic bytes of one cross-attention launch: K_c and V_c of every LIVE (window, head): 2 x 1500 x d bf16 per window.
            # Live windows per launch are counted on the device (wdr_full_get_cross_attn_stats): finished windows exit at once and
            # stream nothing, so with weights that emit EOT a launch serves fewer than B windows.
            live_per_launch = cross_live / cross_launches if cross_launches else float(B)
            per_launch = live_per_launch * 2 * 1500 * dm * 2
            gbs = per_launch * prof["dec_cross"]["records"] / (prof["dec_cross"]["ms"] / 1e3) / 1e9
            kern["dec_cross_attention"] = {"ms_per_step": prof["dec_cross"]["ms"] / args.steps, "launches_per_step": prof["dec_cross"]["records"] / args.steps,
                                           "achieved_gbs": gbs, "frac_hbm": gbs / hbm_peak, "algorithmic_bytes_per_launch": per_launch,
                                           "live_windows_per_launch": live_per_launch, "launches_counted_on_device_last_step": cross_launches}
        for name in ("mel_aux", "layernorm", "decoder", "dec_gemm", "dec_cross_batched", "dtw", "other"):
            if prof[name]["ms"] > 0:
                kern[name] = {"ms_per_step": prof[name]["ms"] / args.steps, "launches_per_step": prof[name]["records"] / args.steps}
        cand = {k: prof[k]["ms"] for k in ("gemm", "attention", "dec_cross")}
        dom = max(cand, key=cand.get)
        if dom == "dec_cross":
            k = kern["dec_cross_attention"]
            roofline = {"bound": "hbm", "kernel": "dec_cross_attn_kernel", "achieved": k["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": k["frac_hbm"], "traffic": None}
        else:
            k = kern[dom]
            roofline = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05)" if dom == "gemm" else "encoder_attention_kernel (tcgen05)",
                        "achieved": k["achieved_tflops"], "peak": tf_peak, "unit": "TFLOP/s", "frac": k["frac"], "traffic": None}
        try:  # per-launch DRAM traffic of that kernel from the committed ncu --set full capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(roofline["kernel"], {})
            if full or roofline["kernel"] != "dec_cross_attn_kernel":
                roofline["traffic"] = tr.get("bytes_per_launch") if args.arch == "large-v3" and B == 120 else None
                roofline["traffic_source"] = tr.get("source")
        except Exception:
            pass
        roofline["peak_source"] = peak_src
        roofline["share_of_step"] = cand[dom] / total_ms
        line = {"metric": metric_name(args.workload), "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload_name(args), "windows_per_gpu": B, "weights": "seeded random-init, bf16 matrices",
                           "pcm": "int16 16 kHz", "decode": ((f"beam search (beam {args.beam})" if args.beam > 1 else "greedy") + (f", temperature ladder +{args.temperature_inc}" if args.temperature_inc > 0 else " T=0") + ", single_segment, token_timestamps, DTW (alignment-head preset)") if full else None,
                           "l2": "no explicit flush: each step streams several GB of activations / KV cache, far above the 126 MB L2"},
                "e2e": {"value": e2e_val, "unit": "audio-s/s", "h2d_bytes_per_step": int(pcm_host.nbytes), "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_s / args.steps * 1e3,
                        "api": ("wdr_full_batch_i16 (pinned host PCM) + whisper.h-style result accessors" if full else
                                "wdr_encode_chunks_i16 (pinned host PCM; result stays in the state) + wdr_state_hidden_digest"), **check},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kern,
                "kernel_ms_sum_per_step": total_ms / args.steps, "profiled_pass_ms_per_step": ms_profiled / args.steps,
                "lanes": int(os.environ.get("WDR_LANES", "1")) if full else 1, "phases_ms_last_step": phases,
                "encoder_tflops_overall": None if full else fl["total_enc"] * B * args.steps / (ms / 1e3) / 1e12}
        if strong is not None:
            line["strong"] = strong
        if not args.no_cpu_baseline and world == 1:  # the CPU arm is reported on rank 0 at N = 1 only
            port = CpuPort(args.arch, args.workload)
            cores = port.threads
            _, t1, _ = port.run(pcm_host, 1)
            n = max(1, min(8, int(15.0 / max(t1, 1e-3)))) if t1 < 15.0 else 0
            if n:
                v, dt, _ = port.run(pcm_host, n)
            else:
                n, v, dt = 1, 30.0 / t1, t1
            line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} of the {B} windows ({dt:.1f} s of CPU work) on {cores} OpenMP threads "
                                              f"({host_cores()} cores in the affinity mask), oracle/wdr_oracle*.c"}
            if full:
                line["parity"] = parity_check(w, st, ctx, port, pcm_host, params, beam_size=max(1, args.beam))
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    st.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
