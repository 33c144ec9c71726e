"""Ad-hoc timing of the sub-flows (development aid, not the bench)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import strugepic_b200 as spic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ppc = int(sys.argv[2]) if len(sys.argv) > 2 else 64
interp = int(sys.argv[3]) if len(sys.argv) > 3 else 0
engine = int(sys.argv[4]) if len(sys.argv) > 4 else 0
print("fp64 probe TFLOP/s:", spic.probe_fp64_tflops(0, 1.0))
s = spic.Simulation((n, n, n), interp=interp, engine=engine)
s.set_uniform_field(0, [0, 0, 0])
s.set_uniform_field(1, [0, 0, 1.0])
t0 = time.time()
s.add_particle_density_uniform(ppc, 100.0, -1.0, 0.01)
s.sync()
npart = s.num_particles()
print("particles", npart, "load s", time.time() - t0)
s.set_option("time_kernels", 1)
for name, fn in (("theta_x", lambda: s.G_Theta(0, 0.25)), ("theta_y", lambda: s.G_Theta(1, 0.25)),
                 ("theta_z", lambda: s.G_Theta(2, 0.25)), ("theta_E", lambda: s.G_Theta_E(0.25)),
                 ("theta_B", lambda: s.G_Theta_B(0.5))):
    fn(); s.sync()
    s.kernel_time_ms(reset=True)
    t0 = time.time()
    for _ in range(3):
        fn()
    s.sync()
    wall = (time.time() - t0) / 3
    ms, nl = s.kernel_time_ms(reset=True)
    print("%-8s wall %.3f ms  particle-kernel %.3f ms/launch  -> %.3e particle-subflows/s" %
          (name, wall * 1e3, ms / max(nl, 1), npart / (ms / max(nl, 1) * 1e-3) if nl else 0))
s.set_option("time_kernels", 0)
s.Theta_map4(0.5); s.sync()
t0 = time.time()
for _ in range(2):
    s.Theta_map4(0.5)
s.sync()
dt = (time.time() - t0) / 2
print("map4: %.3f s/step  %.3e particle-steps/s" % (dt, npart / dt))
print("energy", s.get_total_energy())
