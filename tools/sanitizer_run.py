import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from vqvdb_b200 import BackendType, CodecConfig, DataType, IVQVAECodec, TensorView, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
x = synth.smoke_leaves(n, seed=5)
for prec in ("fp16x2_tc",):
    c = IVQVAECodec.create(CodecConfig(device=CodecConfig.Device.CUDA, encode_precision=prec), BackendType.B200)
    idx = c.encode(TensorView(x, list(x.shape), DataType.FLOAT32)).buffer
    rec = c.decode(TensorView(idx, list(idx.shape), DataType.UINT8)).buffer
    print(prec, idx.sum(), float(rec.sum()))
    c.close()
