// Drop-in include path of the reference (ufo/map/key.h): forwards to the B200-native facade, so a
// caller's #include <ufo/map/key.h> stays as it is when include/ of this repository is on the path.
#pragma once
#include <ufomap_b200/ufomap.hpp>
