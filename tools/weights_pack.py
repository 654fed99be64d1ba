"""Flat weight pack (".vqw") for the B200 VQ-VAE leaf codec.

The reference ships its model only as a TorchScript zip embedded in a C array
(/root/reference/src/Bin/bin_model.h:14, 4 183 532 bytes) and as two ONNX
graphs (src/Bin/bin_onnx.h).  The B200 engine links neither libtorch nor ORT,
so the 45 fp32 tensors of the model's state_dict are serialised once, offline,
into this little-endian flat file that the C++ loader (vqvdb_b200/csrc/
weights.hpp) and the C oracle (oracle/vqvae_oracle.c) both read.

Layout (all little-endian, packed):
    char[8]  magic   "VQVDBW01"
    u32      n_tensors
    u32      in_channels      (1 = float model, 3 = vec3 model)
    u32      embedding_dim    (D, 128)
    u32      num_embeddings   (K, 256)
    repeat n_tensors:
        u32  name_len ; char[name_len] name   (state_dict key, no NUL)
        u32  ndim     ; u32[ndim] dims
        u64  offset   (bytes from start of payload, 64-byte aligned)
        u64  nbytes
    u64      payload_bytes
    <pad to 64-byte file offset>
    payload  fp32 tensors, C-contiguous, in table order

Usage:
    python tools/weights_pack.py float  [--out vqvdb_b200/weights/vqvae_float.vqw]
    python tools/weights_pack.py vec3   [--out vqvdb_b200/weights/vqvae_vec3_seed0.vqw]

`float` needs /root/reference (this container only); `vec3` draws seeded random
weights for the reference's vec3 architecture (python/VQVAE_v2.py:278-325) with numpy
only, because the reference ships no vec3 weights — reproducible anywhere.
"""
from __future__ import annotations

import argparse
import hashlib
import io
import os
import re
import struct
import sys
from collections import OrderedDict

import numpy as np

MAGIC = b"VQVDBW01"
REFERENCE_ROOT = os.environ.get("VQVDB_REFERENCE_ROOT", "/root/reference")
BLOB_SHA256 = "2225fbe5005cccdb8a6c48369c4f26e40b9b7460460428f93720ac6f7c89b2a7"


def write_pack(path: str, tensors: "OrderedDict[str, np.ndarray]", in_channels: int,
               embedding_dim: int, num_embeddings: int) -> str:
    table = io.BytesIO()
    offset = 0
    entries = []
    for name, arr in tensors.items():
        arr = np.ascontiguousarray(arr, dtype="<f4")
        nb = arr.nbytes
        entries.append((name, arr, offset, nb))
        offset = (offset + nb + 63) & ~63
    payload_bytes = offset
    table.write(MAGIC)
    table.write(struct.pack("<IIII", len(entries), in_channels, embedding_dim, num_embeddings))
    for name, arr, off, nb in entries:
        nm = name.encode()
        table.write(struct.pack("<I", len(nm)))
        table.write(nm)
        table.write(struct.pack("<I", arr.ndim))
        table.write(struct.pack("<%dI" % arr.ndim, *arr.shape))
        table.write(struct.pack("<QQ", off, nb))
    table.write(struct.pack("<Q", payload_bytes))
    head = table.getvalue()
    head += b"\0" * ((-len(head)) % 64)
    payload = bytearray(payload_bytes)
    for _, arr, off, nb in entries:
        payload[off:off + nb] = arr.tobytes()
    blob = head + bytes(payload)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(blob)
    return hashlib.sha256(blob).hexdigest()


def read_pack(path: str):
    """Returns (meta dict, OrderedDict name -> np.ndarray[float32])."""
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:8] != MAGIC:
        raise ValueError("%s: not a VQVDBW01 weight pack" % path)
    pos = 8
    n, cin, d, k = struct.unpack_from("<IIII", blob, pos)
    pos += 16
    entries = []
    for _ in range(n):
        (ln,) = struct.unpack_from("<I", blob, pos); pos += 4
        name = blob[pos:pos + ln].decode(); pos += ln
        (nd,) = struct.unpack_from("<I", blob, pos); pos += 4
        dims = struct.unpack_from("<%dI" % nd, blob, pos); pos += 4 * nd
        off, nb = struct.unpack_from("<QQ", blob, pos); pos += 16
        entries.append((name, dims, off, nb))
    (payload_bytes,) = struct.unpack_from("<Q", blob, pos); pos += 8
    pos = (pos + 63) & ~63
    if len(blob) != pos + payload_bytes:
        raise ValueError("%s: truncated weight pack" % path)
    out = OrderedDict()
    for name, dims, off, nb in entries:
        out[name] = np.frombuffer(blob, dtype="<f4", count=nb // 4, offset=pos + off).reshape(dims).copy()
    meta = dict(in_channels=cin, embedding_dim=d, num_embeddings=k,
                sha256=hashlib.sha256(blob).hexdigest())
    return meta, out


def extract_reference_blob() -> bytes:
    """Pulls the TorchScript zip out of the reference's C array (bin_model.h:14)."""
    hdr = os.path.join(REFERENCE_ROOT, "src", "Bin", "bin_model.h")
    src = open(hdr).read()
    i = src.index("g_model_data[g_model_data_size] = {")
    body = src[i:]
    body = body[body.index("{") + 1: body.index("}")]
    blob = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", body))
    got = hashlib.sha256(blob).hexdigest()
    if got != BLOB_SHA256:
        raise RuntimeError("reference model blob hash changed: %s" % got)
    return blob


def load_reference_module():
    import torch
    return torch.jit.load(io.BytesIO(extract_reference_blob()), map_location="cpu").eval()


def vec3_state_dict(seed: int = 0, embedding_dim: int = 128, num_embeddings: int = 256) -> "OrderedDict[str, np.ndarray]":
    """Seeded random weights for the reference's vec3 architecture (EncoderVec3 / DecoderVec3,
    python/VQVAE_v2.py:278-325).  The reference ships no vec3 weights, so any fixed
    weights of that architecture serve; they are drawn with numpy only, so the pack can be regenerated
    bit-identically anywhere (including the GPU box, which has no reference tree).  Names and shapes are
    exactly those of VQVAE(3, D, K).state_dict()."""
    rng = np.random.default_rng(seed)
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def conv(name, cout, cin, k):
        fan_in = cin * k ** 3
        sd[name + ".weight"] = (rng.standard_normal((cout, cin, k, k, k)) * np.sqrt(2.0 / fan_in)).astype("<f4")
        sd[name + ".bias"] = (rng.standard_normal(cout) * 0.05).astype("<f4")

    def gn(name, c):
        sd[name + ".weight"] = (1.0 + 0.1 * rng.standard_normal(c)).astype("<f4")
        sd[name + ".bias"] = (0.1 * rng.standard_normal(c)).astype("<f4")

    def res(name, c):
        gn(name + ".gn1", c)
        conv(name + ".conv1", c, c, 3)
        gn(name + ".gn2", c)
        conv(name + ".conv2", c, c, 3)

    def attn(name, c, r=4):
        sd[name + ".fc.0.weight"] = (rng.standard_normal((c // r, c)) * np.sqrt(2.0 / c)).astype("<f4")
        sd[name + ".fc.2.weight"] = (rng.standard_normal((c, c // r)) * np.sqrt(2.0 / (c // r))).astype("<f4")

    conv("encoder.pre.0", 64, 3, 3)
    gn("encoder.pre.1", 64)
    res("encoder.pre.3", 64)
    conv("encoder.down1", 128, 64, 3)
    res("encoder.res_stack.0", 128)
    res("encoder.res_stack.1", 128)
    attn("encoder.attn", 128)
    conv("encoder.proj", embedding_dim, 128, 1)
    emb = rng.standard_normal((num_embeddings, embedding_dim))
    sd["quantizer.embedding"] = (emb / np.linalg.norm(emb, axis=1, keepdims=True) * 4.0).astype("<f4")
    conv("decoder.stem.0", 128, embedding_dim, 3)
    gn("decoder.stem.1", 128)
    res("decoder.res_stack.0", 128)
    res("decoder.res_stack.1", 128)
    attn("decoder.attn", 128)
    conv("decoder.up_conv", 256, 128, 3)
    conv("decoder.final", 3, 32, 3)
    return sd


def load_vec3_reference_module(pack_path: str):
    """The reference's own VQVAE(3, D, K) (python/VQVAE_v2.py:328-345, imported, never copied) carrying the pack's
    weights — the oracle for config 4.  Needs /root/reference."""
    import torch
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, "python"))
    import VQVAE_v2  # noqa: E402
    meta, tensors = read_pack(pack_path)
    m = VQVAE_v2.VQVAE(3, meta["embedding_dim"], meta["num_embeddings"], 0.25).eval()
    missing, unexpected = m.load_state_dict({k: torch.from_numpy(v) for k, v in tensors.items()}, strict=False)
    assert not unexpected and all(k.startswith("quantizer.") for k in missing), (missing, unexpected)
    return m


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["float", "vec3"])
    ap.add_argument("--out", default=None)
    args = ap.parse_args(argv)
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if args.which == "float":
        mod = load_reference_module()
        cin = 1
        out = args.out or os.path.join(here, "vqvdb_b200", "weights", "vqvae_float.vqw")
    else:
        out = args.out or os.path.join(here, "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw")
        keep = vec3_state_dict(0)
        emb = keep["quantizer.embedding"]
        sha = write_pack(out, keep, 3, emb.shape[1], emb.shape[0])
        print("%s  %d tensors  sha256=%s" % (out, len(keep), sha))
        return
    sd = mod.state_dict()
    keep = OrderedDict()
    for k, v in sd.items():
        if k.startswith("quantizer.") and k != "quantizer.embedding":
            continue  # EMA training buffers (VQVAE_v2.py:104-105) are not inference state
        keep[k] = v.detach().cpu().numpy().astype("<f4")
    emb = keep["quantizer.embedding"]
    sha = write_pack(out, keep, cin, emb.shape[1], emb.shape[0])
    print("%s  %d tensors  sha256=%s" % (out, len(keep), sha))


if __name__ == "__main__":
    main()
