"""A/B timing of the particle-kernel variants (development aid): python scripts/ab_kernels.py [cells] [ppc] [interp]"""
import sys

sys.path.insert(0, ".")
import strugepic_b200 as spic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ppc = int(sys.argv[2]) if len(sys.argv) > 2 else 64
interp = int(sys.argv[3]) if len(sys.argv) > 3 else 0
s = spic.Simulation((n, n, n), interp=interp)
s.set_uniform_field(0, [0, 0, 0])
s.set_uniform_field(1, [0, 0, 1.0])
s.add_particle_density_uniform(ppc, 100.0, -1.0, 0.01)
s.sync()
npart = s.num_particles()
print("fp64 probe TFLOP/s:", spic.probe_fp64_tflops(0, 0.5), "particles", npart)
s.set_option("time_kernels", 1)
s.set_option("fuse", 0)
for variant in (() if "fusedonly" in sys.argv else (2, 3)):
    s.set_option("axis_kernel", variant)
    s.set_option("pushve_kernel", variant)
    for cpb in (32, 64):
        s.set_option("cells_per_block", cpb)
        for _ in range(2):
            s.Theta_map2(0.5)
        s.kernel_times(reset=True)
        for _ in range(2):
            s.Theta_map2(0.5)
        kt = s.kernel_times(reset=True)
        ax, pv = kt["theta_axis"], kt["push_V_E"]
        print("variant %d cpb %3d: theta_axis %.3f ms/launch (%.2f TF)  push_V_E %.3f ms/launch (%.2f TF)" %
              (variant, cpb, ax[0] / ax[1], 718 * npart / (ax[0] / ax[1] * 1e-3) / 1e12 if interp == 0 else 0,
               pv[0] / pv[1], 842 * npart / (pv[0] / pv[1] * 1e-3) / 1e12 if interp == 0 else 0))
s.set_option("fuse", 1)
for bk, cpb in ((2, 64),):
    for _ in range(2):
        s.Theta_map2(0.5)
    s.kernel_times(reset=True)
    for _ in range(2):
        s.Theta_map2(0.5)
    kt = s.kernel_times(reset=True)
    ab, pv, ot = kt["axis_block"], kt["push_V_E"], kt["other"]
    print("fused bk=%d cpb %3d: axis_block %.3f ms/launch = %.3f ms per reference sub-flow (%.2f TF)  push_V_E %.3f ms/launch x %d  other %.3f ms x %d" %
          (bk, cpb, ab[0] / ab[1], ab[0] / ab[1] / 6, 6 * 718 * npart / (ab[0] / ab[1] * 1e-3) / 1e12 if interp == 0 else 0,
           pv[0] / pv[1], pv[1], ot[0] / max(ot[1], 1), ot[1]))
print("energy", s.get_total_energy())
