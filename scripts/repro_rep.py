import sys, os
sys.path[:0] = ['.', 'tests', 'aztec-2.0_b200/python']
import numpy as np, bbg, inputs
from oracle import pyoracle as po
orc = po.Oracle()
pts = orc.read_transcript_g1(inputs.SRS_MINI_POINTS, inputs.SRS_MINI_DIR)
n = 300
rep = np.zeros((n, 8), dtype=np.uint64); rep[:] = pts[7]
sc = inputs.fr_elements(320, n)
res = bbg.msm_points(sc, rep)
print(orc.jac_to_buffer(res).hex())
print(orc.jac_to_buffer(orc.pippenger(sc, rep, stride=1)).hex())
