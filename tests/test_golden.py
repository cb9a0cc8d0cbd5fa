"""Known-answer fixtures (tests/golden/vectors.json, made by tests/golden/make_golden.py).

CPU: the oracle reproduces every vector -- the values the reference's own tests state, and the frozen outputs of earlier
oracle builds.  GPU: the CUDA path through the C ABI gives the same answers."""
import hashlib
import json
import os

import pytest

from tests.oracle_binding import ESCAPED_EXCEPTION, Z_DATA_ERROR

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "vectors.json")) as f:
    GOLDEN = json.load(f)


def _data(v):
    if v.get("data") is not None:
        return bytes.fromhex(v["data"])
    if v.get("gen") == "bytes(70000)":
        return bytes(70000)
    n = v["n"]
    return bytes((i * 7 + 3) & 0xFF for i in range(n))


def test_golden_file_has_every_section():
    assert len(GOLDEN["adler32"]) >= 14 and len(GOLDEN["inflate"]) >= 18 and len(GOLDEN["reference_deflate"]) >= 32
    assert any(v["source"] == "reference-test" for v in GOLDEN["inflate"])


@pytest.mark.parametrize("v", GOLDEN["adler32"], ids=lambda v: "%08x" % v["adler"])
def test_oracle_adler_golden(oracle, v):
    assert oracle.adler32(_data(v)) == v["adler"]


@pytest.mark.parametrize("v", GOLDEN["inflate"], ids=lambda v: v["name"])
def test_oracle_inflate_golden(oracle, v):
    rc, out, second, consumed = oracle.inflate(bytes.fromhex(v["stream"]), cap=1 << 20)
    assert (rc, out.hex(), second, consumed) == (v["rc"], v["out"], v["second"], v["consumed"])


@pytest.mark.parametrize("v", GOLDEN["reference_deflate"], ids=lambda v: "%s-L%d" % (v["name"], v["level"]))
def test_oracle_reference_deflate_golden(oracle, v):
    data = _data(v)
    rc, stream, used = oracle.deflate(data, v["level"])
    assert (rc, used, len(stream)) == (v["rc"], v["bytes_in"], v["stream_len"])
    assert hashlib.sha256(stream).hexdigest() == v["stream_sha256"]
    if v["stream"] is not None:
        assert stream.hex() == v["stream"]


@pytest.mark.gpu
def test_gpu_adler_golden(engine):
    for v in GOLDEN["adler32"]:
        assert engine.adler32(_data(v)) == v["adler"], v["cite"]


@pytest.mark.gpu
@pytest.mark.parametrize("v", GOLDEN["inflate"], ids=lambda v: v["name"])
def test_gpu_inflate_golden(engine, v):
    rc, out, consumed = engine.inflate(bytes.fromhex(v["stream"]), 1 << 20)
    want_rc = Z_DATA_ERROR if v["rc"] == ESCAPED_EXCEPTION else v["rc"]
    assert rc == want_rc
    assert bytes(out).hex() == v["out"]
    if v["rc"] == 1:
        assert consumed == v["consumed"]


@pytest.mark.gpu
def test_gpu_inflates_reference_streams_like_the_reference(engine, oracle):
    """Gate 2: the streams the reference's compressor makes (frozen bytes), decoded on the GPU, give what the reference
    inflate gives -- bytes and return code, including the streams the reference gets wrong."""
    for v in GOLDEN["reference_deflate"]:
        data = _data(v)
        stream = bytes.fromhex(v["stream"]) if v["stream"] is not None else oracle.deflate(data, v["level"])[1]
        rc, out, _ = engine.inflate(stream, len(data) + 4096)
        want_rc = Z_DATA_ERROR if v["inflate_rc"] == ESCAPED_EXCEPTION else v["inflate_rc"]
        assert rc == want_rc, (v["name"], v["level"])
        assert len(out) == v["inflate_out_len"], (v["name"], v["level"])
        if v["inflate_matches_input"]:
            assert bytes(out) == data, (v["name"], v["level"])
