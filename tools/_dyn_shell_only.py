import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
import bench_dynamic as B
from giraffe_b200 import meshes as M
s = M.shell_plate(1000, 500)
B.run("1M Shell_1", s, M.shell_plate_displacements(s), steps=int(sys.argv[1]) if len(sys.argv)>1 else 10)
