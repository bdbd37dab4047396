// oracle/ref_shim/precomp.h -- TEST INFRASTRUCTURE, not product code.
//
// Headless stand-in for /root/reference/template/precomp.h so that the reference's
// hot-path translation units (Intersectable.cpp, lights.cpp, reflection.cpp,
// trianglemesh.cpp + every hot-path header) compile UNMODIFIED on Linux/g++.
// The upstream precomp.h hard-includes <windows.h>, <io.h>, GLFW, glad, OpenCL and zlib
// (template/precomp.h:20,41,45-57) which do not exist here.  Nothing is copied from the
// reference: the pieces of upstream precomp.h that the path needs (typedefs, vector
// structs, Timer, RNG declarations, math, mat4) are cut out of the file WHERE IT LIES at
// build time by oracle/Makefile (sed line ranges -> oracle/_ref/gen/precomp_excerpt.h,
// a git-ignored build artefact) and textually included below.
#pragma once

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "lib/stb_image.h"   // resolved through -I/root/reference (texture.h needs stbi_loadf)
#include <immintrin.h>

using namespace std;          // upstream leaks std (template/precomp.h:33); the headers rely on it

// glibc already declares ::fminf/::fmaxf; upstream re-defines them as `a<b?a:b` / `a>b?a:b`
// (template/precomp.h:364-365).  Renaming keeps exactly those NaN-asymmetric semantics.
#define fminf ref_fminf
#define fmaxf ref_fmaxf

namespace Tmpl8 {}

#include "precomp_excerpt.h"  // generated: oracle/_ref/gen/precomp_excerpt.h

#include "common.h"           // /root/reference/template/common.h, unmodified
