"""Per-step device times of the headline workload (development aid): python scripts/step_times.py [cells] [steps]"""
import sys

import torch

sys.path.insert(0, ".")
import strugepic_b200 as spic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
s = spic.Simulation((n, n, n), interp=0)
s.set_uniform_field(0, [0, 0, 0])
s.set_uniform_field(1, [0, 0, 1.0])
s.add_particle_density_uniform(64, 100.0, -1.0, 0.01)
s.sync()
stream = torch.cuda.ExternalStream(s.stream())
s.set_option("time_kernels", 1)
for k in range(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.kernel_times(reset=True)
    l0 = s.launch_count()
    e0.record(stream)
    s.map(4, 0.5)
    e1.record(stream)
    s.sync()
    kt = s.kernel_times(reset=True)
    print("step %d: %.1f ms, launches %d, kernels %s" % (k, e0.elapsed_time(e1), s.launch_count() - l0,
                                                        {a: round(b[0], 1) for a, b in kt.items()}), flush=True)
