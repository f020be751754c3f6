import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparkfm_b200 import Handle, synth
row_ptr, idx, val, y = synth.regression_c2()
hd = Handle(100_000, 16, task=0, reg=(0.0, 1e-4, 1e-3), step_size=0.02, mini_batch_fraction=0.1)
hd.init_model(0.0, 0.01, 1); hd.load_dataset(row_ptr, idx, val, y)
for it in range(1, 11): hd.train_step(it)
for rep in range(3):
    hd.synchronize(); hd.timer_start(); t0=time.perf_counter(); hd.train(11+100*rep, 100); ms=hd.timer_stop(); print("rep", rep, "event ms/step", ms/100, "wall", (time.perf_counter()-t0)*10)
