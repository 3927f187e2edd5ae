import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import interp_ref as R
from pyg2p_b200 import device, synth
lat,lon,det=synth.octahedral_grid(320)
got=device.latlon_to_xyz(lon,lat,det['radius']).cpu().numpy()
want=R.stack_xyz(*R.to_3d(lon,lat,det['radius']))
print('source O320 xyz bit-equal fraction per component', (got==want).mean(axis=0), 'all three', (got==want).all(axis=1).mean())
tl,tlo=synth.efas_target()
got=device.latlon_to_xyz(tlo,tl,det['radius']).cpu().numpy()
want=R.stack_xyz(*R.to_3d(tlo,tl,det['radius']))
print('EFAS f64 targets', (got==want).mean(axis=0), (got==want).all(axis=1).mean())
