"""Golden vectors the reference itself holds for the hot path."""
import numpy as np

# G1: commented 3x3 KM example, src/km.cpp:237-260 → pairs (0,0),(2,1),(1,2), energy 12
G1_W = np.array([[-5, -2, -100], [-4, -2, -6], [-100, -1, -7]], dtype=np.float64)
# G2: img/GH-ICPworkflow.jpg panels (e),(f): 7 source x 6 target M_cd, T_cd = 30, E_min = 106
G2_CD = np.array([[11, 19, 4, 40, 10, 31], [17, 10, 16, 39, 17, 36], [20, 42, 5, 28, 11, 29],
                  [50, 21, 32, 24, 47, 32], [18, 26, 6, 7, 12, 38], [23, 36, 27, 35, 48, 30],
                  [22, 24, 7, 21, 13, 46]], dtype=np.float64)
G2_PAIRS = [(0, 0), (1, 1), (6, 2), (4, 3), (2, 4)]
