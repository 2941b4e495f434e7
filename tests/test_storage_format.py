"""On-disk format of the reduced operators: directories written by the REAL reference's
OnlineAffineExpansionStorage.save() load into rbnics_b200.online.AffineExpansionStorage, and what ours saves is
file-for-file what the reference wrote (text files byte-identical, .npy files equal arrays)."""
import filecmp
import os

import numpy as np
import pytest

from rbnics_b200.online import AffineExpansionStorage

HERE = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------------------------------------------------
# Directories written by the REAL reference: OnlineAffineExpansionStorage.save() of the numpy backend executed by
# tests/golden/make_reference_golden.py (rbnics/backends/online/basic/affine_expansion_storage.py:56-153), including a
# multi-component (u, s, p) storage and the reference's own component-dict slices of it (slice_to_array.py:27-45).
# ----------------------------------------------------------------------------------------------------------------------
REAL = os.path.join(HERE, "golden", "storage_real")
REAL_STORAGES = {"operator_a": (3,), "operator_f": (2,), "error_estimation_ff": (2, 2), "stokes_block": (2, 1),
                 "int_sized_matrices": (2,), "int_sized_vectors": (2,)}


@pytest.mark.parametrize("name", sorted(REAL_STORAGES))
def test_round_trip_of_reference_written_directories(tmp_path, name):
    st = AffineExpansionStorage(*REAL_STORAGES[name])
    assert st.load(REAL, name) is True
    st.save(tmp_path, name)
    ours, ref = os.path.join(tmp_path, name), os.path.join(REAL, name)
    assert sorted(os.listdir(ours)) == sorted(os.listdir(ref))
    for fn in os.listdir(ref):
        if fn.endswith(".txt"):
            assert filecmp.cmp(os.path.join(ours, fn), os.path.join(ref, fn), shallow=False), fn
        else:
            assert np.array_equal(np.load(os.path.join(ours, fn)), np.load(os.path.join(ref, fn))), fn


def test_component_dict_slicing_equals_the_reference():
    from rbnics_b200.online import OnlineSizeDict
    st = AffineExpansionStorage(2, 1)
    st.load(REAL, "stokes_block")
    assert st._component_name_to_basis_component_length[0] == {"u": 3, "s": 3, "p": 2}
    N = OnlineSizeDict()
    N["u"], N["s"], N["p"] = 2, 1, 2
    sl = st[:N, :N]
    for i in range(2):
        assert np.array_equal(sl[i, 0], np.load(os.path.join(REAL, "stokes_block_sliced_%d.npy" % i)))
    assert sl._component_name_to_basis_component_length == (N, N)
    assert st[:N, :N] is sl                                          # cached like the reference's precomputed slices
    with pytest.raises(ValueError):
        st[:OnlineSizeDict(u=4, s=1, p=1), :N]                       # more basis functions than the component holds


def test_text_files_are_parsed_without_eval(tmp_path):
    """a reduced-model directory must not be able to run code."""
    d = tmp_path / "evil"
    d.mkdir()
    (d / "content_item_type.txt").write_text("__import__('os').system('true')")
    (d / "content_item_shape.txt").write_text("None")
    with pytest.raises(ValueError):
        AffineExpansionStorage(1).load(tmp_path, "evil")
