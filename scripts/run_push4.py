"""A few Theta_E with the pair-blocked push kernel at 128^3 x 64 ppc (ncu target)."""
import sys

sys.path.insert(0, ".")
import strugepic_b200 as spic

s = spic.Simulation((128, 128, 128), interp=0)
s.set_uniform_field(0, [0.1, 0.2, 0.3])
s.set_uniform_field(1, [0, 0, 1.0])
s.add_particle_density_uniform(64, 100.0, -1.0, 0.01)
s.Theta_map2(0.5)
s.set_option("pushve_kernel", int(sys.argv[1]) if len(sys.argv) > 1 else 4)
for _ in range(3):
    s.G_Theta_E(0.1)
s.sync()
