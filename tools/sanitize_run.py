import numpy as np, sys
sys.path.insert(0, '.')
from dpdfnet_b200.engine import Engine
from dpdfnet_b200.spec import get_spec
from dpdfnet_b200.weights import random_checkpoint
name = sys.argv[1] if len(sys.argv) > 1 else "dpdfnet2"
spec = get_spec(name)
B, T = 70, 3
eng = Engine(spec, random_checkpoint(spec, 0), max_streams=B)
for k in ("intra_tc", "post_tc", "sep_tc", "gru_tc", "dft_tc"):
    eng.set_option(k, 1)
eng.set_option("graph", 0)
for extra in sys.argv[2:]:
    k, v = extra.split("=")
    eng.set_option(k, int(v))
pcm = (np.random.default_rng(0).standard_normal((B, T * spec.hop)) * 0.1).astype(np.float32)
out = eng.run_pcm_host(pcm)
print(name, sys.argv[2:], "finite", np.isfinite(out).all(), float(np.abs(out).max()))
eng.close()
