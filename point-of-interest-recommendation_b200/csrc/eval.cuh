// eval.cuh -- scoring + top-K kernels
#pragma once
#include "common.cuh"
