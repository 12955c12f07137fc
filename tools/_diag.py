import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from giraffe_b200 import capi, meshes as M
from oracle.portdrv import PortOracle
import util
m = M.shell_plate(250, 120, warp=0.01, gravity=(0.0, 0.0, -9.81))
d = M.shell_plate_displacements(m)
port = PortOracle(threads=4).load(m)
port.assemble(d)
pp = port.vectors()
os.environ["GFA_RING"] = "0"
c = capi.Assembler(m).set_dofs(); c.assemble(d); cv = c.vectors()
os.environ["GFA_RING"] = "1"
r = capi.Assembler(m).set_dofs(); r.assemble(d); rv = r.vectors()
print(r.pipeline_info())
for name, a, b in (("port-classic", pp[0], cv[0]), ("port-ring", pp[0], rv[0]), ("classic-ring", cv[0], rv[0])):
    diff = np.abs(a - b); i = np.argsort(diff)[-5:]
    print(name, "max|v|", np.abs(a).max(), "max diff", diff.max(), "n>1e-9", int((diff > 1e-9 * np.abs(a).max()).sum()), "idx", i, "a", a[i], "b", b[i])
gls, nf, nx = M.number_dofs(m)
bad = np.nonzero(np.abs(pp[0] - rv[0]) > 1e-10 * np.abs(pp[0]).max())[0]
if len(bad):
    # nodes of the bad DOFs
    inv = {}
    g = gls.reshape(-1)
    pos = np.nonzero(g > 0)[0]; node_of = np.zeros(nf, int); node_of[g[pos] - 1] = pos // 6
    nodes = np.unique(node_of[bad]); print("bad dofs", len(bad), "bad nodes", len(nodes), nodes[:20], "n_nodes", m.n_nodes)
    print("xyz of bad nodes", m.xyz[nodes[:10]])
