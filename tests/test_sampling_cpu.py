"""CPU checks of the sampling restatements (oracle/d2_ref.py): the Philox generator against the Random123 known-answer
vector, and detectron2's sampler - replayed with the reference's own torch.randperm draws - against the outputs of the
reference's unmodified label_and_sample_proposals / label_and_sample_anchors (tests/golden/labels_ref.pt)."""
import numpy as np
import torch

from conftest import load_golden
from oracle import d2_ref

LABELS = load_golden("labels_ref.pt")


def test_philox4x32_10_known_answer():
    # Random123 kat_vectors: philox4x32-10, counter 0 0 0 0, key 0 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    assert int(d2_ref.philox_keys(np.array([0]), 0, 0, 0)[0]) == 0x6627E8D5
    # counter ffffffff x4, key ffffffff x2 -> 408f276d 41c83b0e a20bc7c6 6d5451fd
    assert int(d2_ref.philox_keys(np.array([0xFFFFFFFF]), 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF)[0]) == 0x408F276D


def test_sample_proposals_replays_the_reference_run():
    for c in LABELS["roi"]:
        torch.manual_seed(c["torch_seed"])
        gt = c["gt_classes_cat"]
        cls = gt[c["matched_idxs"]].clone() if gt.numel() else torch.zeros_like(c["matched_idxs"]) + c["num_classes"]
        if gt.numel():
            cls[c["matched_labels"] == 0] = c["num_classes"]
            cls[c["matched_labels"] == -1] = -1
        n_pos = int(((cls != -1) & (cls != c["num_classes"])).sum())
        n_neg = int((cls == c["num_classes"]).sum())
        perms = (torch.randperm(n_pos), torch.randperm(n_neg))
        sampled, temp = d2_ref.sample_proposals(c["matched_idxs"], c["matched_labels"], gt, c["num_classes"],
                                                c["batch_size_per_image"], c["positive_fraction"], perms)
        bg = temp == c["num_classes"]
        assert torch.equal(temp[bg], c["sampled"]["bg"]["gt_classes"]), c["label"]
        len_a = len(c["a"]["gt_boxes"])
        m = c["matched_idxs"][sampled]
        mask_a = (m < len_a) & ~bg
        assert torch.equal(c["a"]["gt_boxes"][m[mask_a]], c["sampled"]["a"]["gt_boxes"]), c["label"]


def test_device_policy_draw_is_a_uniform_subset_in_key_order():
    labels = torch.tensor([1, 0, -1, 0, 1, 1, 0, 0, 0, 1] * 50)
    pos, neg = d2_ref.subsample_labels(labels, 64, 0.25, 0, seed=7, offset=3)
    assert len(pos) == 16 and len(neg) == 48
    assert bool((labels[pos] == 1).all()) and bool((labels[neg] == 0).all())
    assert len(set(pos.tolist())) == 16 and len(set(neg.tolist())) == 48
    kp = d2_ref.philox_keys(pos.numpy(), 0, 7, 3)
    assert bool((kp[:-1] <= kp[1:]).all())
    pos2, _ = d2_ref.subsample_labels(labels, 64, 0.25, 0, seed=7, offset=4)
    assert pos.tolist() != pos2.tolist()
