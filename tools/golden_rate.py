import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyg2p_b200 import _lib as L, device
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
for case in ['o24_box_f64', 'o24_laea_f32mv', 'regional_global_f64', 'o24_coincident_f64']:
    g = np.load(os.path.join(G, f'build_{case}.npz'))
    ix = device.SourceIndex(g['src_lon'], g['src_lat'], float(g['radius']))
    mub = ix.min_upper_bound(int(g['nj']))
    for mode, k in (('nearest', 1), ('invdist', 4)):
        out = ix.knn(g['tgt_lon'], g['tgt_lat'], k=k, mode=L.MODE_NEAREST if k == 1 else L.MODE_INVDIST, mub=mub)
        f32 = g['tgt_lat'].dtype == np.float32
        tbl = device.DeviceTable(out['idx'], out['coef'], g['src_lon'].size, coef_f32=(f32 and k == 1))
        idx, w = tbl.indexes_coeffs()
        gi, gw = g[f'{mode}_indexes'], g[f'{mode}_weights']
        same_i = (idx.reshape(len(idx), -1) == gi.reshape(len(gi), -1)).all(axis=1)
        same_w = ((w.reshape(len(w), -1) == gw.reshape(len(gw), -1)) | (np.isnan(w.reshape(len(w), -1)) & np.isnan(gw.reshape(len(gw), -1)))).all(axis=1)
        print(f'{case:24s} {mode:8s} mub equal {mub == float(g[mode + "_mub"])}  rows: idx identical {same_i.mean():.4f}  '
              f'coeffs bit-identical {same_w.mean():.4f}  both {np.mean(same_i & same_w):.4f}')
