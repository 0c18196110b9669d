// geoie.cuh -- GeoIE train step kernels (filled in below api_more.cuh)
#pragma once
#include "common.cuh"
