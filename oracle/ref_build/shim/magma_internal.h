// intentionally empty: the reference includes this MAGMA-internal header but uses nothing from it on the CPU path
