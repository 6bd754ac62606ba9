// intentionally empty: the reference GPU kernels include this MAGMA header but use nothing from it
