#ifndef RUB_SHIM_VOLK_MALLOC_H
#define RUB_SHIM_VOLK_MALLOC_H
#include "volk.h"
#endif
