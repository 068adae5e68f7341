"""Device-side input pipeline (csrc/scene_prep.cu) against a numpy restatement of the reference's per-object processing
(data/datasets/base.py:697-741) fed with the SAME sample indices, plus the distributional properties of the sampling and of
the token / object masking (data/data_utils.py:76-121)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def ragged_scene(B, O, seed, big=True):
    rng = np.random.default_rng(seed)
    counts = np.zeros(B * O, np.int64)
    for b in range(B):
        n_obj = rng.integers(3, O + 1)
        counts[b * O: b * O + n_obj] = rng.integers(5, 4096 if big else 600, size=n_obj)
    counts[1] = 1024                      # exactly P points: a permutation of the object
    counts[2] = 1                         # a single point: max norm 0 -> scale 1
    off = np.concatenate([[0], np.cumsum(counts)])
    raw = rng.standard_normal((off[-1], 6)).astype(np.float32)
    raw[:, :3] = raw[:, :3] * rng.uniform(0.1, 2.0) + rng.uniform(-4, 4, size=(1, 3)).astype(np.float32)
    return raw, off, counts


def test_objects_match_reference_processing_given_the_same_indices():
    from sceneverse_b200 import input_pipeline as ip
    B, O, P = 4, 20, 1024
    raw, off, counts = ragged_scene(B, O, 3)
    fts, locs, masks, idx = ip.prepare_objects(torch.from_numpy(raw).cuda(), torch.from_numpy(off).cuda(), B, O, P, seed=11,
                                               return_indices=True)
    fts, locs, masks, idx = fts.cpu().numpy().reshape(B * O, P, 6), locs.cpu().numpy().reshape(B * O, 6), \
        masks.cpu().numpy().reshape(-1), idx.cpu().numpy().reshape(B * O, P)
    for s in range(B * O):
        n = counts[s]
        if n == 0:                         # dataset_wrapper.py:62-72: pad = 1.0 points, 0.0 locs, mask False
            assert not masks[s] and (fts[s] == 1.0).all() and (locs[s] == 0).all() and (idx[s] == -1).all()
            continue
        pcd = raw[off[s]: off[s + 1]].astype(np.float64)
        assert masks[s]
        np.testing.assert_allclose(locs[s, :3], pcd[:, :3].mean(0), rtol=1e-5, atol=1e-5)       # base.py:710-712
        np.testing.assert_allclose(locs[s, 3:], pcd[:, :3].max(0) - pcd[:, :3].min(0), rtol=1e-6, atol=1e-6)
        ii = idx[s]
        assert ii.min() >= 0 and ii.max() < n
        if n >= P:                         # np.random.choice(n, P, replace=False): P distinct indices
            assert len(np.unique(ii)) == P
        samp = pcd[ii]                     # base.py:721-729 on the same subsample
        samp[:, :3] -= samp[:, :3].mean(0)
        md = np.sqrt((samp[:, :3] ** 2).sum(1)).max()
        samp[:, :3] /= (1.0 if md < 1e-6 else md)
        np.testing.assert_allclose(fts[s], samp, rtol=2e-5, atol=2e-5)
        if n > 1:
            assert abs(np.sqrt((fts[s][:, :3] ** 2).sum(1)).max() - 1.0) < 1e-5


def test_sampling_is_uniform_and_reproducible():
    from sceneverse_b200 import input_pipeline as ip
    B, O, P = 1, 8, 1024
    n = 3000
    counts = np.full(B * O, n, np.int64)
    off = np.concatenate([[0], np.cumsum(counts)])
    raw = np.random.default_rng(0).standard_normal((off[-1], 6)).astype(np.float32)
    rawc, offc = torch.from_numpy(raw).cuda(), torch.from_numpy(off).cuda()
    hits = np.zeros(n)
    for seed in range(40):
        idx = ip.prepare_objects(rawc, offc, B, O, P, seed=seed, return_indices=True)[3].cpu().numpy().reshape(-1, P)
        for row in idx:
            assert len(np.unique(row)) == P
            hits[row] += 1
    # 320 draws of 1024 / 3000: every index is chosen ~109 times (binomial sigma ~8.5)
    assert abs(hits.mean() - 320 * P / n) < 1e-9 and hits.min() > 60 and hits.max() < 165 and hits.std() < 12
    a = ip.prepare_objects(rawc, offc, B, O, P, seed=5, return_indices=True)[3]
    b = ip.prepare_objects(rawc, offc, B, O, P, seed=5, return_indices=True)[3]
    c = ip.prepare_objects(rawc, offc, B, O, P, seed=6, return_indices=True)[3]
    assert torch.equal(a, b) and not torch.equal(a, c)
    # fewer points than P: sampling with replacement covers the object
    small = np.full(B * O, 100, np.int64)
    offs = np.concatenate([[0], np.cumsum(small)])
    idx = ip.prepare_objects(rawc[:offs[-1]].contiguous(), torch.from_numpy(offs).cuda(), B, O, P, seed=1, return_indices=True)[3]
    assert idx.min() >= 0 and idx.max() < 100 and len(torch.unique(idx[0, 0])) > 95


def test_token_and_object_masking_statistics():
    from sceneverse_b200 import input_pipeline as ip
    B, L = 512, 50
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(1000, 30000, (B, L), generator=g).cuda()
    lens = torch.randint(8, L + 1, (B,), generator=g)
    am = (torch.arange(L)[None, :] < lens[:, None]).long().cuda()
    out, lab = ip.mask_tokens(ids, am, 0.15, seed=9)
    valid = am.bool()
    assert (lab[~valid] == -1).all() and (out[~valid] == ids[~valid]).all()         # padding untouched, never supervised
    sup = lab != -1
    assert (lab[sup] == ids[sup]).all()                                              # label = the original token
    frac = sup[valid].float().mean().item()
    assert abs(frac - 0.15) < 0.01
    masked = (out == 103) & sup
    kept = (out == ids) & sup
    assert abs(masked.sum().item() / sup.sum().item() - 0.8) < 0.03 and abs(kept.sum().item() / sup.sum().item() - 0.1) < 0.03
    assert (out[~sup] == ids[~sup]).all()
    om = torch.rand(64, 80, generator=g).cuda() > 0.3
    sem = ip.mask_objects(om, 0.1, seed=3)
    assert not (sem & ~om).any() and abs(sem.sum().item() / om.sum().item() - 0.9) < 0.03


def test_built_data_dict_feeds_the_model():
    """The dict built on the device has the keys / dtypes / padding of dataset_wrapper.py:38-111 and runs through the encoders."""
    from sceneverse_b200 import input_pipeline as ip
    B, O, P = 2, 16, 1024
    raw, off, counts = ragged_scene(B, O, 8)
    labels = torch.where(torch.from_numpy(counts.reshape(B, O)) > 0, torch.randint(0, 607, (B, O)), torch.tensor(-100)).cuda()
    ids = torch.randint(1000, 30000, (B, 50)).cuda()
    am = torch.ones(B, 50, dtype=torch.int64).cuda()
    d = ip.build_data_dict(torch.from_numpy(raw).cuda(), torch.from_numpy(off).cuda(), labels, ids, am,
                           torch.zeros(B, 1, dtype=torch.int64).cuda(), B, O, P, seed=4)
    assert d["obj_fts"].shape == (B, O, P, 6) and d["obj_fts"].dtype == torch.float32 and d["obj_masks"].dtype == torch.bool
    assert d["obj_locs"].shape == (B, O, 6) and d["masked_lm_labels"].dtype == torch.int64 and d["obj_sem_masks"].dtype == torch.bool
    from sceneverse_b200 import weights
    from sceneverse_b200.modules import vision
    enc = vision.PointOpenVocabEncoder(None, freeze=True, text_features=weights.synthetic_tensor("text_features", (607, 768))).cuda().eval()
    with torch.no_grad():
        obj, pre, sem = enc(d["obj_fts"], d["obj_locs"], d["obj_masks"], d["obj_sem_masks"], d["obj_labels"], 1, 1)
    assert torch.isfinite(obj).all() and obj.shape == (B, O, 768)
