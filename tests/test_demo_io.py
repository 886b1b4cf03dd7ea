"""CPU: the demo.py output contract (per-sample npy/txt, --allinone packing) byte for byte against a literal
restatement of the reference's writer (demo.py:176-214; the script itself needs omegaconf / pytorch_lightning
and cannot be imported here)."""
import glob
import os

import numpy as np
import pytest
import torch

from mld_b200 import demo_io


def _reference_writer(output_dir, task, text, length, rep_joints, outall):
    """demo.py:176-214, verbatim control flow (variable names as in the reference)."""
    rep_lst, texts_lst = [], []
    batch = {"length": length, "text": text}
    for joints in rep_joints:
        nsample = len(joints)
        id = 0
        for i in range(nsample):
            npypath = str(os.path.join(output_dir, f"{task}_{length[i]}_batch{id}_{i}.npy"))
            with open(npypath.replace(".npy", ".txt"), "w") as text_file:
                text_file.write(batch["text"][i])
            np.save(npypath, joints[i].detach().cpu().numpy())
        if outall:
            rep_lst.append(joints)
            texts_lst.append(batch["text"])
    if outall:
        grouped_lst = []
        for n in range(nsample):
            grouped_lst.append(torch.cat([r[n][None] for r in rep_lst], dim=0)[None])
        combinedOut = torch.cat(grouped_lst, dim=0)
        npypath = str(os.path.join(output_dir, f"{task}_{length[i]}_all.npy"))
        np.save(npypath, combinedOut.detach().cpu().numpy())
        with open(npypath.replace('npy', 'txt'), "w") as text_file:
            for texts in texts_lst:
                for t in texts:
                    text_file.write(t)
                    text_file.write('\n')


def _tree(d):
    return {os.path.basename(p): open(p, "rb").read() for p in sorted(glob.glob(os.path.join(d, "*")))}


@pytest.mark.parametrize("lengths,outall", [([196, 64, 120], False), ([88, 88], True)])
def test_files_match_reference_writer(tmp_path, lengths, outall):
    g = torch.Generator().manual_seed(7)
    texts = [f"a person does thing {i}" for i in range(len(lengths))]
    reps = [[torch.randn(n, 22, 3, generator=g) for n in lengths] for _ in range(3)]
    a, b = tmp_path / "ours", tmp_path / "ref"
    a.mkdir(), b.mkdir()
    rep_lst, texts_lst = [], []
    for joints in reps:
        demo_io.write_samples(str(a), "Example", texts, lengths, joints)
        rep_lst.append(joints)
        texts_lst.append(texts)
    if outall:
        p = demo_io.write_allinone(str(a), "Example", lengths, rep_lst, texts_lst)
        assert np.load(p).shape == (len(lengths), 3, lengths[0], 22, 3)       # [n_samples, n_rep, nframes, 22, 3]
    _reference_writer(str(b), "Example", texts, lengths, reps, outall)
    ta, tb = _tree(str(a)), _tree(str(b))
    assert ta.keys() == tb.keys()
    assert all(ta[k] == tb[k] for k in ta), "file contents differ from the reference writer"
    one = np.load(os.path.join(str(a), f"Example_{lengths[0]}_batch0_0.npy"))
    assert one.shape == (lengths[0], 22, 3) and one.dtype == np.float32


def test_allinone_rejects_ragged_lengths(tmp_path):
    reps = [[torch.zeros(10, 22, 3), torch.zeros(12, 22, 3)]]
    with pytest.raises(ValueError, match="Lengths of motions are different"):
        demo_io.write_allinone(str(tmp_path), "Example", [10, 12], reps, [["a", "b"]])
