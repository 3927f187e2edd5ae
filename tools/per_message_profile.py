#!/usr/bin/env python
"""Where the time of one blocking `Interpolator.interpolate_scipy(pageable MaskedArray)` call goes (O1280 -> EFAS, mask
policy propagate): cProfile over 300 calls."""
import cProfile
import os
import pstats
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from bench import _Ctx  # noqa: E402
from pyg2p_b200 import synth  # noqa: E402
from pyg2p_b200.interpolator import Interpolator  # noqa: E402
from pyg2p_b200.maps import LatLong  # noqa: E402

lat, lon, det = synth.octahedral_grid(1280)
tl, tlo = synth.efas_target()
tmp = tempfile.mkdtemp()
ctx = _Ctx({'interpolation.mode': 'invdist', 'interpolation.create': True, 'interpolation.parallel': True,
            'interpolation.dirs': {'user': tmp, 'global': tmp}, 'input.file': 'x.grib',
            'interpolation.latMap': LatLong(tl, tlo, mv=synth.NC_FILL_F64, identifier='efas'), 'interpolation.lonMap': None})
it = Interpolator(ctx, synth.GRIB_MV, mask_policy='propagate')
mask = synth.missing_mask(lat)
fields = [np.ma.masked_array(synth.field_2t(lat, lon, member=i), mask=mask) for i in range(4)]
gid = '0.0$359.929$0$2560$6599680$reduced_gg'
for i in range(8):
    it.interpolate_scipy(lat, lon, fields[i % 4], gid, det)
pr = cProfile.Profile()
pr.enable()
for i in range(300):
    it.interpolate_scipy(lat, lon, fields[i % 4], gid, det)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(18)
