import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import megakv_b200 as mk
from oracle import pyoracle as po
from tests import helpers as H
from tests.test_gpu_parity import gpu_insert
for seed in (1, 2, 3):
    rng = np.random.default_rng(seed)
    mem_p = 20
    slots = (1 << mem_p) // 8
    iel = H.random_requests(rng, int(0.9 * slots))
    parts = np.array_split(iel, 8)
    env = {"to_b2": [], "displaced": [], "dropped": []}
    for k in range(4):
        o = po.Oracle(mem_p, po.CUCKOO)
        for part in parts:
            o.insert(part[rng.permutation(len(part))])
        w = o.stats.as_dict()
        for key in env: env[key].append(w[key])
    for nparts in (8, 64, 512):
        t = mk.DeviceTable(mem_p, po.CUCKOO, mk.LAYOUT_PAIRS); st = mk.DeviceStats()
        for part in np.array_split(iel, nparts): gpu_insert(t, part, stats=st)
        s = st.read()
        print(seed, nparts, {k: (min(v), max(v)) for k, v in env.items()}, {"to_b2": s["ins_to_b2"], "displaced": s["ins_displaced"], "dropped": s["ins_dropped"], "retry": s["ins_cas_retry"]}, flush=True)
