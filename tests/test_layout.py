"""Plan layout: Python table == C table; pack/unpack round trip; parameter counts of SURVEY 8(a)."""
import torch

from shapley_vit_b200 import _lib, layout, synth


def test_param_counts_match_survey():
    assert layout.num_params(layout.vit_preset("tiny", image=32, n_cls=10)) == 5_489_482
    assert layout.num_params(layout.vit_preset("small", image=224, n_cls=10)) == 21_669_514
    assert layout.num_params(layout.vit_preset("base", image=224, n_cls=10)) == 85_806_346
    assert layout.num_params(layout.vit_preset("large", image=224, n_cls=10)) == 303_311_882


def test_flops_match_survey():
    cfg = layout.vit_preset("base", image=224, n_cls=10)
    assert abs(cfg.flops_per_image() / 1e9 - 35.13) < 0.05
    assert abs(layout.vit_preset("tiny", image=32, n_cls=10).flops_per_image() / 1e9 - 0.055) < 0.001


def test_c_layout_equals_python_layout():
    for name, image in (("tiny", 32), ("small", 224), ("base", 224), ("large", 224)):
        cfg = layout.vit_preset(name, image=image, n_cls=10)
        vs, ms, segs = _lib.layout_segments(cfg)
        lay = layout.plan_layout(cfg)
        assert (vs, ms) == (lay.vec_size, lay.mat_size)
        assert len(segs) == len(lay.segments)
        for a, b in zip(segs, lay.segments):
            assert (a.kind, a.layer, a.region, a.offset, a.size) == (b.kind, b.layer, b.region, b.offset, b.size)
            assert a.rows * a.cols == b.size
            assert a.offset % 64 == 0


def test_qkv_segments_are_contiguous():
    lay = layout.plan_layout(layout.vit_preset("tiny", image=32, n_cls=10))
    h = lay.cfg.hidden
    for l in range(lay.cfg.layers):
        q, k, v = (lay.find(kind, l) for kind in (layout.K_WQ, layout.K_WK, layout.K_WV))
        assert k.offset == q.offset + h * h and v.offset == k.offset + h * h
        bq, bk, bv = (lay.find(kind, l) for kind in (layout.K_BQ, layout.K_BK, layout.K_BV))
        assert bk.offset == bq.offset + h and bv.offset == bk.offset + h


def test_pack_unpack_roundtrip():
    cfg = layout.vit_preset("tiny", image=32, n_cls=10, layers=2)
    sd = synth.make_state_dict(cfg, 5)
    lay = layout.plan_layout(cfg)
    row = layout.pack_state_dict(lay, sd)
    assert row.numel() == lay.total and lay.total % 64 == 0
    back = layout.unpack_row(lay, row)
    assert list(back.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(back[k], sd[k])
    assert float(row.sum()) != 0.0


def test_module_state_dict_keys_are_hf_keys():
    from shapley_vit_b200.models.vit import ViTForImageClassification, infer_config

    cfg = layout.vit_preset("tiny", image=32, n_cls=10, layers=3)
    m = ViTForImageClassification(cfg)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == layout.state_dict_spec(cfg)
    assert infer_config(m.state_dict(), heads=3) == cfg
    prefixed = {"module.base_model.model." + k: v for k, v in m.state_dict().items()}
    assert infer_config(prefixed, heads=3) == cfg
