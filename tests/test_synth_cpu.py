"""CPU: the package's synthetic generators equal the oracle's (both benchmark arms see the same data)."""
import torch

from oracle import decap as o_decap
from oracle import dinov2 as o_vit
from oracle import pipeline as o_pipe


def test_generators_match_oracle():
    from patchioner_b200 import synth

    a, b = synth.make_vit_weights(5), o_vit.make_weights(5)
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    a, b = synth.make_decoder_weights(5), o_decap.make_weights(5)
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(synth.synth_images(2, 28, 3), o_pipe.synth_images(2, 28, 3))
    assert torch.equal(synth.synth_boxes(3, 7, 224, 2, pad="set"), o_pipe.synth_boxes(3, 7, 224, 2, pad="set"))
    assert synth.synth_traces(3, 2) == o_pipe.synth_traces(3, 2)
    assert torch.equal(synth.synth_bank(500, 64, 1), o_pipe.synth_bank(500, 64, 1))


def test_viecap_synth_weights_equal_oracle_generator():
    from oracle import viecap as ov
    from patchioner_b200 import synth

    a = synth.make_viecap_weights(seed=4321, n_layer_gpt=2, n_layer_map=2)
    b = ov.make_weights(seed=4321, n_layer_gpt=2, n_layer_map=2)
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    t1, t2 = synth.WordTokenizer(), ov.ToyTokenizer()
    s = "There are person, traffic light in image."
    assert t1.encode(s) == t2.encode(s) and t1.decode(t1.encode(s)) == s
