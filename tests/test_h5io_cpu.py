"""CPU: the package's HDF5 writer / reader (orbithunter_b200/h5io.py; reference core.py:1855-1950, io.py:25-139) against the
independent minimal reader of the test infrastructure (oracle/h5mini.py, written from the format notes in SURVEY.md App. A)
and, in the build container, against the reference's own h5py-written fixture."""
import os

import numpy as np
import pytest

from oracle import refshim
from oracle.h5mini import H5Mini
from orbithunter_b200 import h5io


def _leaf(i, cls="OrbitKS", shape=(3, 4)):
    return (np.arange(np.prod(shape), dtype=float).reshape(shape) + i,
            {"basis": "modes", "class": cls, "discretization": np.array([4, 6]), "parameters": np.array([1.5 + i, 2.5, 0.25]), "cost": 1e-12 * i})


def test_written_file_is_read_by_the_independent_reader(tmp_path):
    f = str(tmp_path / "t.h5")
    tree = {"rpo": {"0": _leaf(0, "RelativeOrbitKS")}, "family": {str(i): _leaf(i) for i in range(300)}}
    tree["rpo"]["0"][1]["frame"] = "comoving"
    h5io.write_tree(f, tree)
    got = H5Mini(f).datasets()
    assert len(got) == 301
    arr, attrs = got["rpo/0"]
    assert np.array_equal(arr, tree["rpo"]["0"][0]) and attrs["frame"] == "comoving" and attrs["class"] == "RelativeOrbitKS"
    assert attrs["basis"] == "modes" and list(attrs["discretization"]) == [4, 6] and np.array_equal(attrs["parameters"], [1.5, 2.5, 0.25])
    for i in (0, 7, 123, 299):
        arr, attrs = got[f"family/{i}"]
        assert np.array_equal(arr, tree["family"][str(i)][0]) and attrs["parameters"][0] == 1.5 + i and attrs["cost"] == 1e-12 * i
    back = h5io.h5_tree(f)
    assert set(back) == {"rpo", "family"} and set(back["family"]) == set(tree["family"])
    assert np.array_equal(back["family"]["42"][0], tree["family"]["42"][0]) and back["rpo"]["0"][1]["frame"] == "comoving"
    # structure: superblock version 0 with 8-byte offsets, end-of-file address = file size, groups sorted by name
    raw = open(f, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    assert int.from_bytes(raw[40:48], "little") == len(raw)
    assert list(H5Mini(f)._group_entries(H5Mini(f).root_btree, H5Mini(f).root_heap)) == ["family", "rpo"]


def test_name_collisions_follow_the_reference_rule(tmp_path):
    """core.py:1902-1918: numeric dataset names count up, other names get _0, _1, ... appended; groups nest on '/'."""
    tree = {}
    for expect in ("0", "1", "2"):
        parts = h5io._free_name(tree, "", "")
        assert parts == [expect]
        h5io._insert(tree, parts, _leaf(0))
    for expect in ("fam/a/t44p000", "fam/a/t44p000_0", "fam/a/t44p000_1"):
        parts = h5io._free_name(tree, "fam/a", "t44p000")
        assert "/".join(parts) == expect
        h5io._insert(tree, parts, _leaf(1))
    f = str(tmp_path / "c.h5")
    h5io.write_tree(f, tree)
    back = h5io.h5_tree(f)
    assert set(back) == {"0", "1", "2", "fam"} and set(back["fam"]["a"]) == {"t44p000", "t44p000_0", "t44p000_1"}


@pytest.mark.skipif(not refshim.available(), reason="needs the reference tree (build container)")
def test_reads_the_reference_fixture_written_by_h5py():
    path = os.path.join(refshim.REFERENCE_ROOT, "tests", "test_data.h5")
    if not os.path.exists(path):
        pytest.skip("fixture not carried along")
    ours, mini = h5io.h5_tree(path), H5Mini(path).datasets()
    assert len(ours) == 7
    for key, (arr, attrs) in mini.items():
        group, name = key.split("/")
        a2, at2 = ours[group][name]
        assert np.array_equal(arr, a2) and at2["class"] == attrs["class"] and at2["basis"] == attrs["basis"]
        assert np.array_equal(at2["parameters"], attrs["parameters"]) and list(at2["discretization"]) == list(attrs["discretization"])


def test_symbol_strings_follow_the_reference_grouping():
    """io.to_symbol_string / to_symbol_array (io.py:172-189); expected strings are outputs of the reference."""
    from orbithunter_b200.h5io import to_symbol_array, to_symbol_string

    cases = [(np.array([[0, 1], [2, 0]]), "01_20"), (np.arange(8).reshape(2, 2, 2) % 3, "01_20__12_01"), (np.array([1, 2, 0]), "120"),
             (np.arange(24).reshape(2, 3, 4) % 5, "012_340_123_401__234_012_340_123"), (np.arange(6).reshape(3, 2), "01_23_45")]
    for arr, want in cases:
        assert to_symbol_string(arr) == want
        assert np.array_equal(to_symbol_array(want, arr.shape), arr)


def test_read_tileset_builds_the_tiling_dictionary(tmp_path):
    """io.read_tileset (io.py:142-169): {key: orbit} in the order of the names given."""
    f = str(tmp_path / "tiles.h5")
    tree = {"streak": {"0": _leaf(0)}, "defect": {"0": _leaf(1, "RelativeOrbitKS")}, "wiggle": {"0": _leaf(2, "AntisymmetricOrbitKS", (3, 2))}}
    h5io.write_tree(f, tree)
    tiles = h5io.read_tileset(f, (0, 1, 2), ("streak", "wiggle", "defect"))
    assert [type(tiles[k]).__name__ for k in (0, 1, 2)] == ["OrbitKS", "AntisymmetricOrbitKS", "RelativeOrbitKS"]
    assert np.array_equal(tiles[1].state, tree["wiggle"]["0"][0]) and tiles[2].parameters == (2.5, 2.5, 0.25)
    with pytest.raises(AssertionError):
        h5io.read_tileset(f, (0, 1), ("streak",))
