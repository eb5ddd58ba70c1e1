def to_inference_data(*a, **k):
    raise NotImplementedError
get_samples = to_xarray = ess_bulk = rhat = to_inference_data
