"""Small encode + decode through the default (tcgen05) paths for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitizer_run.py 300
    compute-sanitizer --tool racecheck python tools/sanitizer_run.py 300

Also drives the encoder near-tie path on every row (debug tap stage 5).  n <= 296 runs the decoder with one tile per CTA, larger n with two."""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np  # noqa: E402,F401
import torch  # noqa: E402

from vqvdb_b200 import BackendType, CodecConfig, DataType, IVQVAECodec, TensorView, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
x = synth.smoke_leaves(n, seed=5)
for dec in ("default",):
    c = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, encode_precision="fp16x2_tc", decode_precision=dec), BackendType.B200)
    idx = c.encode(TensorView(x, list(x.shape), DataType.FLOAT32)).buffer
    rec = c.decode(TensorView(idx, list(idx.shape), DataType.UINT8)).buffer
    print(c.encode_path, c.decode_path, idx.sum(), float(rec.sum()))
    if dec == "default":  # exact (near-tie) path of the VQ on every row
        xd = torch.from_numpy(x).cuda()
        tap = torch.zeros((n, 128, 64), dtype=torch.float32, device="cuda")
        idx_d = torch.empty((n, 4, 4, 4), dtype=torch.uint8, device="cuda")
        c.debug_encode_tap(xd, n, 5, tap, idx_d, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        print("forced exact path: indices equal", bool((idx_d.cpu().numpy() == idx).all()))
    c.close()

# the vec3 model's 128-channel tensor-core decoder (decode_tc128.cu), odd leaf count (a spare leaf slot in the last group)
pack = os.path.join(os.getcwd(), "vqvdb_b200", "weights", "vqvae_vec3_seed0.vqw")
if os.path.exists(pack):
    c = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, source=pack), BackendType.B200)
    m = min(n, 75)
    vidx = np.random.default_rng(3).integers(0, 256, size=(m, 4, 4, 4), dtype=np.uint8)
    vrec = c.decode(TensorView(vidx, list(vidx.shape), DataType.UINT8)).buffer
    print(c.decode_path, float(vrec.sum()))
    # ... and its tensor-core encoder (encode_tc128_front.cu + encode_tc128.cu), same odd count
    vx = synth.smoke_leaves(m, seed=9, channels=3, sparse=True)
    venc = c.encode(TensorView(vx, list(vx.shape), DataType.FLOAT32)).buffer
    print(c.encode_path, int(venc.astype(np.int64).sum()))
    # more leaves than CTAs (from a CTA's second leaf on, pre.0 is computed ahead under the previous leaf's MMAs), and the
    # codebook search's exact path forced on every row (debug tap stage 3)
    m2 = min(n, 301)
    vx2 = synth.smoke_leaves(m2, seed=10, channels=3)
    venc2 = c.encode(TensorView(vx2, list(vx2.shape), DataType.FLOAT32)).buffer
    xd = torch.from_numpy(vx2).cuda()
    tap = torch.zeros((m2, 128, 64), dtype=torch.float32, device="cuda")
    idx_d = torch.empty((m2, 4, 4, 4), dtype=torch.uint8, device="cuda")
    c.debug_encode_tap(xd, m2, 3, tap, idx_d, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print("vec3, %d leaves, forced exact path: indices equal" % m2, bool((idx_d.cpu().numpy() == venc2).all()))
    c.close()
