"""sha256 of the weights after a few deterministic train steps on Digital-Music -- for comparing kernel variants that
must be BIT-identical (e.g. `TLSAN_SORT_IMPL=count python tools/state_hash.py` vs `python tools/state_hash.py`)."""
import hashlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tlsan_oracle as O          # fixtures + collate only
from tests.util import load_digital_music, model_from_params

dm = load_digital_music()
cfg = O.default_config(*dm.counts)
params = O.randomize_params(O.init_params(cfg, seed=1234), seed=7)
m = model_from_params(params, dm.icl, cfg)
for step in range(5):
    n = (32, 301, 64, 512, 7)[step]
    m.train(None, O.collate_train(dm.train_set[step * 600:step * 600 + n], 10), 1.0)
h = hashlib.sha256()
for k, v in m.state_dict().items():
    h.update(np.ascontiguousarray(v.numpy()).tobytes())
print({k: os.environ.get(k) for k in ("TLSAN_SORT_IMPL", "TLSAN_BWD_LONG", "TLSAN_FUSED_IMPL")}, h.hexdigest())
