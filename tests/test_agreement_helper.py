"""The end-to-end agreement report (tests/common.py:e2e_agreement) on stand-in engine results, without a GPU: a perfect copy of
the oracle's detections gives no exceptions, a dropped detection is listed with its margins, and a dropped detection that is NOT
margin-limited is flagged as unexplained (the gate the `-m gpu` end-to-end tests assert on)."""
import types

import numpy as np
import torch

from tests.common import synthetic_heads, post_config, e2e_agreement


def _fake_padded(ref, drop=()):
    keep_rows = [i for i in range(len(ref['pred'])) if i not in drop]
    k = len(keep_rows)
    cand_pred = torch.from_numpy(ref['cand']['pred'].astype(np.int32))[None]
    # position of every kept detection inside the candidate list
    pos = {(int(p), int(c)): i for i, (p, c) in enumerate(zip(ref['cand']['pred'], ref['cand']['cls']))}
    keep = torch.tensor([[pos[(int(ref['pred'][i]), int(ref['cls'][i]))] for i in keep_rows]], dtype=torch.int32)
    return types.SimpleNamespace(
        count=torch.tensor([k], dtype=torch.int32), keep=keep, candidates={'pred': cand_pred},
        cls=torch.from_numpy(ref['cls'][keep_rows])[None], det=torch.from_numpy(ref['bbox'][keep_rows])[None],
        mask=torch.from_numpy(ref['mask'][keep_rows].astype(np.uint8))[None])


def test_agreement_report_lists_and_explains_exceptions():
    from oracle.post_oracle import PostProcessOracle
    cfg = post_config(64, 96, 0.005)
    oracle = PostProcessOracle(cfg['grid_size'], cfg['image_size'], cfg['anchors'], cfg['anchor_mask'], 80, conf_thresh=0.005)
    heads = synthetic_heads(1, 64, 96, seed=4)
    ref = oracle([(b.numpy(), o.numpy()) for b, o in heads])[0]
    assert len(ref['pred']) > 10
    rep = e2e_agreement(ref, _fake_padded(ref), 0)
    assert rep['exceptions'] == [] and rep['matched'] == rep['reference_detections'] and rep['min_mask_iou'] == 1.0
    assert rep['max_box_err'] == 0.0 and rep['unexplained'] == 0
    # drop the highest-scoring detection: far from every cut and from every NMS threshold -> a real disagreement
    top = int(np.argmax(ref['bbox'][:, 4]))
    rep = e2e_agreement(ref, _fake_padded(ref, drop=(top,)), 0)
    assert len(rep['exceptions']) == 1 and rep['exceptions'][0]['kept_by'] == 'reference'
    assert rep['exceptions'][0]['pred'] == int(ref['pred'][top]) and rep['matched'] == rep['reference_detections'] - 1
    # with a noise bound as wide as the score range everything is "margin-limited": the flag follows the bound
    wide = e2e_agreement(ref, _fake_padded(ref, drop=(top,)), 0, score_noise=1.0)
    assert wide['unexplained'] == 0
