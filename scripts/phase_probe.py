import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparkfm_b200 import Handle, synth
which = sys.argv[1]
if which == "c2":
    row_ptr, idx, val, y = synth.regression_c2()
    hd = Handle(100_000, 16, task=0, reg=(0.0, 1e-4, 1e-3), step_size=0.02, mini_batch_fraction=0.1)
    hd.init_model(0.0, 0.01, 1); hd.load_dataset(row_ptr, idx, val, y)
else:
    card = synth.ctr_field_log2_cards(24); cdf, off = synth.zipf_tables(card)
    hd = Handle(10_000_000, 64, task=1, reg=(0.0, 0.0, 1e-5), step_size=0.1, mini_batch_fraction=500_000/8_000_000)
    hd.init_model(0.0, 0.01, 1); hd.synth_ctr_dataset(8_000_000, 0, card, cdf, off, 20260104)
hd.train(1, 5)
hd.set_phase_timing(True); hd.stats_reset(); hd.train(6, 10); st = hd.stats(); hd.set_phase_timing(False)
print(which, {k: round(st[k]/10, 4) for k in ("ms_forward","ms_sort","ms_reduce","ms_update")}, "rows/step", st["train_rows"]/10, "launches/step", st["kernel_launches"]/10)
hd.synchronize(); t0=time.perf_counter(); hd.train(16, 20); hd.synchronize(); print("wall ms/step", (time.perf_counter()-t0)/20*1e3)
