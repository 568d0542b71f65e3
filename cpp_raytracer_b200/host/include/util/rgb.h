// Compatibility forwarder: the reference keeps this API in include/util/rgb.h; here the whole
// host-side mirror lives in b200rt/raytracer.hpp (scene description only -- the work is done by libb200rt.so).
#pragma once
#include "../b200rt/raytracer.hpp"
