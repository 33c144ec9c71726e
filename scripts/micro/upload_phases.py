"""Phases of spic_set_particles when a caller re-uploads every step (SPIC_TRACE_PHASES=1 python scripts/micro/upload_phases.py [cells]).
The trace synchronises the stream at every mark: a diagnostic, never a bench number."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import strugepic_b200 as spic

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 128
sim = spic.Simulation((cells,) * 3, interp=0)
sim.add_particle_density_uniform(64, 1.0, -1.0, 0.05)
n = sim.num_particles(0)
host = [torch.empty(n, dtype=torch.float64, pin_memory=True).numpy() for _ in range(6)]
sim.get_particles(0, out=host)
for k in range(3):
    sim.sync()
    t0 = time.perf_counter()
    sim.set_particles(0, *host)
    sim.sync()
    t1 = time.perf_counter()
    sim.map(4, 0.5)
    sim.sync()
    t2 = time.perf_counter()
    sim.get_particles(0, out=host)
    sim.sync()
    t3 = time.perf_counter()
    print("round %d: %d particles, set %.1f ms (%.1f GB/s of list), map %.1f ms, get %.1f ms" %
          (k, n, 1e3 * (t1 - t0), 48e-9 * n / (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)), file=sys.stderr, flush=True)
