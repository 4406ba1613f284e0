from .base import raytrace, single_theta_trace_parallel, calc_weights, calc_weights_parallel, calculate_spherical_ray  # noqa: F401
