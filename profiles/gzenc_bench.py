"""k_gz_encode timing: 'clean' stage with and without TGSF_FLAG_GZ_BLOCKS on synthetic config[1]/config[0] reads."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from tgsfilter_b200 import synth
from tgsfilter_b200.engine import FilterEngine
for cfg, n in ((2, 20000), (1, 20000)):
    batch = synth.make_config(cfg, n, with_names=False)
    for gz in (False, True):
        p = synth.config_params(cfg)
        p.gz_blocks = gz
        with FilterEngine(p) as eng:
            for rep in range(3):
                eng.submit(batch)
                if gz:
                    t0 = time.perf_counter(); blob, spans = eng.collect_gz(); t_gz = time.perf_counter() - t0
                eng.collect()
                st = eng.last_stage_ms()
            extra = " blob %.3f B/base, collect_gz %.1f ms" % (blob.size / batch.n_bases, t_gz * 1e3) if gz else ""
            print("config", cfg, "bases %.3g" % batch.n_bases, "gz" if gz else "  ", "clean stage %.2f ms" % st["clean"], "kernel total %.2f ms" % sum(st.values()), extra)
