"""ONNX-initializer reader: weights read out of encoder.onnx + decoder.onnx without ONNX Runtime
(the reference's OnnxModelPaths / model-directory sources, OnnxBackendFactory.cpp:97-145) must equal the weight pack.

Two sources of graphs:
  * synthesised here from the shipped pack with a ~40-line protobuf writer, in the exporter's conventions (named conv /
    codebook initializers behind "vqvae.", anonymous [C,1,1,1] GroupNorm affine constants consumed by scoped Mul / Add
    nodes, transposed anonymous MatMul weights for the bias-free Linear layers) — runs everywhere;
  * the reference's own embedded graphs (src/Bin/bin_onnx.h), when /root/reference is present (dev container only).
"""
import os
import re
import struct
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tools"))
import weights_pack as wp  # noqa: E402

PACK = os.path.join(REPO, "vqvdb_b200", "weights", "vqvae_float.vqw")
REF_ONNX_HEADER = "/root/reference/src/Bin/bin_onnx.h"


def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _field(num, wire, payload):
    if wire == 0:
        return _varint(num << 3) + _varint(payload)
    return _varint((num << 3) | 2) + _varint(len(payload)) + payload


def _tensor(name, arr):
    arr = np.ascontiguousarray(arr, dtype="<f4")
    body = b"".join(_field(1, 0, d) for d in arr.shape) + _field(2, 0, 1) + _field(8, 2, name.encode()) + _field(9, 2, arr.tobytes())
    return _field(5, 2, body)


def _node(op, name, inputs, outputs):
    body = b"".join(_field(1, 2, i.encode()) for i in inputs) + b"".join(_field(2, 2, o.encode()) for o in outputs)
    body += _field(3, 2, name.encode()) + _field(4, 2, op.encode())
    return _field(1, 2, body)


def synthesise_graph(tensors, prefix):
    """One ModelProto holding the tensors whose names start with `prefix` (plus the codebook), the way torch.onnx
    export writes them."""
    graph, anon = b"", 100
    for name, arr in tensors.items():
        if not (name.startswith(prefix) or name == "quantizer.embedding"):
            continue
        mod, leaf = name.rsplit(".", 1) if "." in name else (name, "")
        scope = "/" + mod.replace(".", "/")
        if arr.ndim == 1 and (".gn" in name or re.search(r"\.(pre|stem)\.1\.", name)):   # GroupNorm affine
            anon += 1
            op = "Mul" if leaf == "weight" else "Add"
            const = "onnx::%s_%d" % (op, anon)
            graph += _tensor(const, arr.reshape(-1, 1, 1, 1))
            graph += _node(op, "%s/%s_2" % (scope, op), [scope + "/x", const], ["%s/%s_2_output_0" % (scope, op)])
        elif arr.ndim == 2 and ".attn.fc." in name:                                       # bias-free Linear
            anon += 1
            const = "onnx::MatMul_%d" % anon
            graph += _tensor(const, arr.T)
            graph += _node("MatMul", scope + "/MatMul", [scope + "/x", const], [scope + "/MatMul_output_0"])
        else:
            graph += _tensor("vqvae." + name, arr)
            if name == "quantizer.embedding":   # the exporter also emits the transposed codebook for the distance GEMM
                graph += _tensor("onnx::MatMul_999", arr.T)
                graph += _node("MatMul", "/MatMul", ["/Reshape_output_0", "onnx::MatMul_999"], ["/MatMul_output_0"])
    return _field(7, 2, graph)


def _check_against_pack(enc_path, dec_path, tmp_path):
    import vqvdb_b200
    out = str(tmp_path / "from_onnx.vqw")
    vqvdb_b200.convert_onnx(enc_path, dec_path, out)
    meta, got = wp.read_pack(out)
    meta0, want = wp.read_pack(PACK)
    assert (meta["in_channels"], meta["embedding_dim"], meta["num_embeddings"]) == (1, 128, 256)
    assert set(got) == set(want), sorted(set(want) ^ set(got))
    for name, arr in want.items():
        assert got[name].shape == arr.shape, name
        assert np.array_equal(got[name], arr), name


def test_reader_on_synthesised_graphs(tmp_path):
    _, tensors = wp.read_pack(PACK)
    enc, dec = tmp_path / "encoder.onnx", tmp_path / "decoder.onnx"
    enc.write_bytes(synthesise_graph(tensors, "encoder."))
    dec.write_bytes(synthesise_graph(tensors, "decoder."))
    _check_against_pack(str(enc), str(dec), tmp_path)


def test_reader_rejects_garbage(tmp_path):
    import vqvdb_b200
    bad = tmp_path / "bad.onnx"
    bad.write_bytes(b"\x3a\xff\xff\xff\xff\x0f not a protobuf")
    with pytest.raises(RuntimeError, match="convert_onnx failed"):
        vqvdb_b200.convert_onnx(str(bad), str(bad), str(tmp_path / "x.vqw"))


@pytest.mark.skipif(not os.path.exists(REF_ONNX_HEADER), reason="reference tree not present (dev container only)")
def test_reader_on_the_reference_graphs(tmp_path):
    src = open(REF_ONNX_HEADER, "r", errors="ignore").read()

    def extract(name):
        i = src.index(name + "[")
        j = src.index("{", i)
        k = src.index("}", j)
        return bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", src[j + 1:k]))
    enc, dec = tmp_path / "encoder.onnx", tmp_path / "decoder.onnx"
    enc.write_bytes(extract("encoder_model_data"))
    dec.write_bytes(extract("decoder_model_data"))
    _check_against_pack(str(enc), str(dec), tmp_path)


@pytest.mark.gpu
def test_codec_from_onnx_sources_matches_embedded_model(tmp_path):
    from vqvdb_b200 import BackendType, CodecConfig, DataType, IVQVAECodec, OnnxModelPaths, TensorView, synth
    _, tensors = wp.read_pack(PACK)
    (tmp_path / "encoder.onnx").write_bytes(synthesise_graph(tensors, "encoder."))
    (tmp_path / "decoder.onnx").write_bytes(synthesise_graph(tensors, "decoder."))
    x = synth.smoke_leaves(64, seed=3)
    base = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA), BackendType.B200)
    want = base.encode(TensorView(x, list(x.shape), DataType.FLOAT32)).buffer
    rec = base.decode(TensorView(want, list(want.shape), DataType.UINT8)).buffer
    base.close()
    for source in (OnnxModelPaths(str(tmp_path / "encoder.onnx"), str(tmp_path / "decoder.onnx")), str(tmp_path)):
        c = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, source=source), BackendType.B200)
        assert c is not None
        assert np.array_equal(c.encode(TensorView(x, list(x.shape), DataType.FLOAT32)).buffer, want)
        assert np.array_equal(c.decode(TensorView(want, list(want.shape), DataType.UINT8)).buffer, rec)
        c.close()
