// intentionally (almost) empty: the reference host layer includes this MAGMA header but uses nothing from it
#include "magma_v2.h"
