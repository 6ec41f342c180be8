class Patch:  # names imported at uav_env.py:11-12
    pass


class Circle:
    pass
