"""state_dict compatibility of the drop-in modules (no GPU): the reference's torch-geometric 1.7.2 parameter names are
canonical; checkpoints with the names of torch-geometric >= 2.0 load as well."""
import pytest
import torch

from hic_gnn_b200 import layers, models

V1_KEYS = ["att_l", "att_r", "bias", "lin_l.weight", "lin_r.weight"]


def test_gatconv_keys_are_the_1_7_2_names():
    conv = layers.GATConv(512, 256, heads=2)
    assert list(conv.state_dict().keys()) == V1_KEYS
    assert conv.lin_r is conv.lin_l


@pytest.mark.parametrize("rename", [
    {"lin_l.weight": "lin_src.weight", "lin_r.weight": "lin_dst.weight", "att_l": "att_src", "att_r": "att_dst"},   # PyG 2.0 - 2.2
    {"lin_l.weight": "lin.weight", "lin_r.weight": None, "att_l": "att_src", "att_r": "att_dst"},                   # PyG >= 2.3
    {"lin_r.weight": None},                                                                                          # only one alias stored
])
def test_newer_pyg_checkpoints_load(rename):
    torch.manual_seed(0)
    src = layers.GATConv(512, 256, heads=2)
    sd = {}
    for k, v in src.state_dict().items():
        new = rename.get(k, k)
        if new is not None:
            sd[new] = v.clone()
    dst = layers.GATConv(512, 256, heads=2)
    res = dst.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k in V1_KEYS:
        assert torch.equal(dst.state_dict()[k], src.state_dict()[k]), k
    # and through a parent module (prefix "conv.")
    net_sd = models.GATNetHeadsChanged3LayersLeakyReLUv2().state_dict()
    renamed = {}
    for k, v in net_sd.items():
        head, _, tail = k.partition(".")
        new = rename.get(tail, tail) if head == "conv" else tail
        if new is not None:
            renamed[f"{head}.{new}" if tail else head] = v
    net = models.GATNetHeadsChanged3LayersLeakyReLUv2()
    assert not net.load_state_dict(renamed, strict=True).missing_keys
    assert all(torch.equal(net.state_dict()[k], net_sd[k]) for k in net_sd)


def test_unknown_keys_are_still_reported():
    conv = layers.GATConv(512, 256, heads=2)
    sd = dict(conv.state_dict(), lin_edge=torch.zeros(1))
    with pytest.raises(RuntimeError, match="lin_edge"):
        conv.load_state_dict(sd, strict=True)
