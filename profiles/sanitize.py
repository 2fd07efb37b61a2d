"""Small RK4 + Euler runs through the DMMA kernels for compute-sanitizer (memcheck / racecheck):
python profiles/sanitize.py [order] [flow: 0 = zero mean flow -> warp-specialised kernel, 1 = mean flow -> tiled kernel] [kernel id, e.g. 4 / 5 = Bernstein]
e.g.  compute-sanitizer --tool memcheck python profiles/sanitize.py 4 0 4"""
import sys
import numpy as np
sys.path.insert(0, "/root/repo")
import __graft_entry__ as g
pkg = g.load_package()
order = int(sys.argv[1]) if len(sys.argv) > 1 else 4
flow = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mesh = pkg.Mesh(pkg.Model.make_cube(3, -10.0, 10.0, order), pkg.Config())
mesh.set_physics(c0=343.0, rho0=1.225, v0=(30.0, 10.0, 0.0) if flow else (0.0, 0.0, 0.0), dt=1e-5)
b = np.nonzero(mesh.fIsBoundary)[0]; mesh.fBC[b[::2]] = 1
eng = pkg.Engine(mesh)
if len(sys.argv) > 3:
    eng.set_option("kernel", int(sys.argv[3]))
eng.set_state(np.random.default_rng(0).standard_normal((4, mesh.N)))
eng.run(pkg.RUNGE_KUTTA, 0.0, 2)
eng.run(pkg.EULER1, 0.0, 1)
print(eng.kernel_name, float(np.abs(eng.get_state()).max()))
