"""getdist.chains placeholder (see package docstring)."""


class WeightedSampleError(Exception):
    pass


class WeightedSamples:  # pragma: no cover - placeholder
    pass


print_load_details = False
