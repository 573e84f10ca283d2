"""CPU checks of the ResNet34 embedding restatement (oracle/resnet.py): layer table, folding, shapes, determinism."""
import numpy as np

from oracle import resnet, weights as W


def test_layer_table():
    specs = resnet.conv_specs()
    assert len(specs) == 36  # conv1 + 16 BasicBlocks x 2 + 3 projection shortcuts
    assert [s[0] for s in specs if s[0].endswith("shortcut")] == ["layer2.0.shortcut", "layer3.0.shortcut", "layer4.0.shortcut"]
    assert specs[0] == ("conv1", 1, 32, 3, 1) and specs[-1] == ("layer4.2.conv2", 256, 256, 3, 1)
    # ~4.5 GFLOP/s of audio per the survey's estimate for 2 s; the exact table gives 45.5 MFLOP per fbank frame at T = 198
    assert abs(resnet.flops(198) / 198 / 1e6 - 45.5) < 1.0


def test_fold_is_bf16_and_matches_batchnorm():
    import torch
    rng = np.random.default_rng(0)
    w = rng.standard_normal((8, 4, 3, 3)).astype(np.float32)
    g, beta, mean = (rng.standard_normal(8).astype(np.float32) for _ in range(3))
    var = rng.random(8).astype(np.float32) + 0.5
    wf, bf = resnet.fold(w, g, beta, mean, var)
    assert np.array_equal(wf, W.bf16_round(wf))
    x = torch.from_numpy(rng.standard_normal((1, 4, 10, 12)).astype(np.float32))
    ref = torch.nn.functional.batch_norm(torch.nn.functional.conv2d(x, torch.from_numpy(w), padding=1), torch.from_numpy(mean), torch.from_numpy(var),
                                         torch.from_numpy(g), torch.from_numpy(beta), False, 0.0, 1e-5)
    got = torch.nn.functional.conv2d(x, torch.from_numpy(wf), torch.from_numpy(bf), padding=1)
    assert float((ref - got).abs().max()) < 0.05 * float(ref.abs().max())  # bf16 rounding of the folded kernel only


def test_forward_shapes_and_determinism(oracle):
    w = resnet.resnet_weights(1234)
    rng = np.random.default_rng(1)
    pcm = (rng.standard_normal(8000) * 2000).astype(np.int16)
    e1 = resnet.compute(pcm, w, oracle.kaldi_fbank)
    e2 = resnet.compute(pcm, resnet.resnet_weights(1234), oracle.kaldi_fbank)
    assert e1.shape == (256,) and e1.dtype == np.float32 and np.isfinite(e1).all()
    assert np.array_equal(e1, e2)
    assert resnet.compute(pcm[:399], w, oracle.kaldi_fbank) is None
    one = resnet.compute(pcm[:400], w, oracle.kaldi_fbank)  # a single frame: std term is sqrt(1e-7), not NaN
    assert np.isfinite(one).all()
