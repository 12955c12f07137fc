import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
os.environ.update(GFA_RING="1", GFA_RING_CHUNK_KB="256", GFA_FUSED_DEBUG="1", GFA_FUSED_TIMEOUT_MS="1000")
from giraffe_b200 import capi, meshes as M
m = M.shell_plate(40, 25, warp=0.01)
d = M.shell_plate_displacements(m)
a = capi.Assembler(m).set_dofs()
print(a.pipeline_info(), flush=True)
try:
    a.assemble(d)
    print("ok", a.timing(), a.launch_count())
except Exception as e:
    print("ERR", e)
