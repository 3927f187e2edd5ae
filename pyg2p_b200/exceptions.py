"""Error codes of the interpolation boundary (same numbers and `get_exc(code, details)` call shape as the
reference's pyg2p/exceptions.py:1-82, restricted to the codes the interpolation path can raise)."""

NOT_EXISTING_MAPS = 1300
NO_INTERTABLE_CREATED = 8100
INVALID_INTERPOL_METHOD = 8400
WEIRD_STUFF = 9000
MALFORMED_GRIB_LATLON = 5000
GRIB_INTERPOLATION_WITHOUT_GID = 6000
NOT_IMPLEMENTED = 6100

_MESSAGES = {
    NOT_EXISTING_MAPS: "Latitude or longitude maps doesn't exist. Check filenames in commands json file.",
    MALFORMED_GRIB_LATLON: 'Interpolating with not existing lat/lons. Probably a geopotential grib.\n'
                           'Geopotentials must be interpolated with an intertable. Try to create it first.',
    GRIB_INTERPOLATION_WITHOUT_GID: 'Trying to interpolate manipulated values with no more reference to original gribs, '
                                    'using grib_api interpolation methods. Interlookup table must be created, first.',
    NOT_IMPLEMENTED: 'Not implemented.',
    NO_INTERTABLE_CREATED: 'Interpolation table was not found and needs to be created '
                           'but -B option was not set on command line.',
    INVALID_INTERPOL_METHOD: 'Interpolation method not valid',
    WEIRD_STUFF: 'Cannot continue: weird stuff happening',
}


class ApplicationException(Exception):
    def __init__(self, inner, code, error):
        super().__init__(str(error))
        self._innerException = inner
        self.code = code
        self.message = str(error)

    @classmethod
    def get_exc(cls, code, details=''):
        return cls(None, code, f'{_MESSAGES.get(code, "error")} - {details}')

    def __str__(self):
        return self.message
