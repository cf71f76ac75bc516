def cameras_points_cartesian_product(*a, **k):
    raise NotImplementedError
