"""Where a host.diarize call spends its time (cProfile over 3 calls after 2 warm-up calls)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import wdr_b200 as w
from hostmirror import host as H
from conftest import synth_audio
pcm = synth_audio(4001, 600.0, n_speakers=4)
seg = w.Segmenter(seed=1234); emb = w.EmbeddingExtractor(seed=1234)
for _ in range(2):
    H.diarize(seg, emb, pcm, 0.5, w.SIZE_MAX, "leader")
for i in range(3):
    t = time.perf_counter(); H.diarize(seg, emb, pcm, 0.5, w.SIZE_MAX, "leader"); print("call", i, (time.perf_counter() - t) * 1e3, "ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    H.diarize(seg, emb, pcm, 0.5, w.SIZE_MAX, "leader")
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
