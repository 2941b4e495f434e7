"""On-disk format of the reduced operators: a directory written by the reference's own io
classes (tests/golden/storage_ref, produced by executing rbnics/utils/io/{text_io,numpy_io}.py in
tests/golden/make_golden.py) loads into rbnics_b200.online.AffineExpansionStorage, and what ours saves is
file-for-file what the reference wrote (text files byte-identical, .npy files equal arrays)."""
import filecmp
import os

import numpy as np

from rbnics_b200.online import AffineExpansionStorage

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "golden", "storage_ref")
G = np.load(os.path.join(HERE, "golden", "reference_sweep_tutorial01.npz"))


def test_load_reference_written_storage():
    Qa, N = G["A"].shape[0], G["A"].shape[1]
    a = AffineExpansionStorage(Qa)
    assert a.load(REF, "operator_a") is True
    assert np.array_equal(a.stacked(), G["A"])
    assert a.load(REF, "operator_a") is False                      # the reference does not load twice either
    f = AffineExpansionStorage(G["F"].shape[0])
    f.load(REF, "operator_f")
    assert np.array_equal(f.stacked(), G["F"])
    ff = AffineExpansionStorage(*G["Gff"].shape)
    ff.load(REF, "error_estimation_ff")
    assert np.array_equal(ff.stacked(), G["Gff"])
    assert np.array_equal(a[:2, :2].stacked(), G["A"][:, :2, :2])  # slicing after load


def test_save_matches_reference_files(tmp_path):
    a = AffineExpansionStorage(tuple(G["A"][q] for q in range(G["A"].shape[0])))
    f = AffineExpansionStorage(tuple(G["F"][q] for q in range(G["F"].shape[0])))
    ff = AffineExpansionStorage(*G["Gff"].shape)
    for i in range(G["Gff"].shape[0]):
        for j in range(G["Gff"].shape[1]):
            ff[i, j] = float(G["Gff"][i, j])
    for name, st in (("operator_a", a), ("operator_f", f), ("error_estimation_ff", ff)):
        st.save(tmp_path, name)
        ours, ref = os.path.join(tmp_path, name), os.path.join(REF, name)
        assert sorted(os.listdir(ours)) == sorted(os.listdir(ref))
        for fn in os.listdir(ref):
            if fn.endswith(".txt"):
                assert filecmp.cmp(os.path.join(ours, fn), os.path.join(ref, fn), shallow=False), fn
            else:
                assert np.array_equal(np.load(os.path.join(ours, fn)), np.load(os.path.join(ref, fn))), fn
