"""Property tests (hypothesis) for the host-only native parsers: whatever bytes a FASTA or an
eq_classes file holds, the C packers must agree with the reference's Python readers."""
import os

import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from karma_b200 import fasta
from oracle import readgraph_oracle as ro
from tests.test_fasta import reference_read_fasta

ALPHABET = "ACGTNacgt>  \t_|=0123456789xyz"
line_endings = st.sampled_from(["\n", "\r\n", "\r"])
lines = st.lists(st.tuples(st.text(alphabet=ALPHABET, min_size=0, max_size=40), line_endings), min_size=0, max_size=25)


@pytest.fixture(params=["one_chunk", "tiny_chunks"])
def chunking(request, monkeypatch):
    """The packer cuts the file image into per-thread chunks of >= 4 MB; `tiny_chunks` forces a cut at
    (almost) every line so that records, sequence lines and "\r\n" pairs straddle chunk borders."""
    if request.param == "tiny_chunks":
        monkeypatch.setenv("KB_FASTA_MIN_CHUNK", "1")
        monkeypatch.setenv("KB_EQ_MIN_CHUNK", "1")
        monkeypatch.setenv("KB_HOST_THREADS", "13")
    return request.param


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(body=lines, trailing=st.booleans())
def test_fasta_packer_equals_reference_reader(tmp_path, chunking, body, trailing):
    text = "".join(l + e for l, e in body)
    if not trailing and body:
        text = text[:-len(body[-1][1])]
    path = os.path.join(str(tmp_path), "f.fa")
    with open(path, "wb") as f:
        f.write(text.encode("ascii"))
    want = reference_read_fasta(path)
    got = fasta.read_fasta_file(path)
    assert list(got.items()) == list(want.items())
    pk = got.packed()
    if pk is not None:
        bases, offsets, key_len = pk
        assert [bases[offsets[i]:offsets[i + 1]].tobytes().decode() for i in range(len(want))] == list(want.values())
        assert key_len.tolist() == [len(k) for k in want]


classes_st = st.lists(st.tuples(st.sampled_from(["1", "2", "3", "7"]),
                                st.lists(st.integers(0, 11), min_size=1, max_size=6, unique=True),
                                st.integers(0, 500)), min_size=0, max_size=40)


@settings(max_examples=100, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(classes=classes_st)
def test_eq_parser_equals_python_parser(tmp_path, chunking, classes):
    from karma_b200 import read_graph as rg
    names = ["contig_%d x" % i for i in range(12)]
    path = os.path.join(str(tmp_path), "eq.txt")
    ro.write_eq_file(path, names, classes)
    p = rg.parse(path)
    want_names, want_classes = ro.parse_eq_file(path)
    assert p["names"] == want_names
    off = p["class_off"].tolist()
    assert [(("1" if p["skip"][i] else "x"), p["ids"][off[i]:off[i + 1]].tolist(), int(p["counts"][i])) for i in range(len(want_classes))] == \
           [(("1" if f == "1" else "x"), ids, c) for f, ids, c in want_classes]
