import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyg2p_b200 import device, synth
lat, lon, det = synth.octahedral_grid(1280)
d_lon, d_lat = torch.from_numpy(lon).cuda(), torch.from_numpy(lat).cuda()
for r in range(4):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    x = device.latlon_to_xyz(d_lon, d_lat, det['radius'])
    b.record()
    torch.cuda.synchronize()
    print('xyz kernel + alloc', a.elapsed_time(b), 'ms')
