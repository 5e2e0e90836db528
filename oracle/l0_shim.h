/* oracle/l0_shim.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Force-included (-include) in front of the UNMODIFIED reference sources so that they compile
 * with g++/libstdc++ on Linux and so that their RNG becomes thread-private and seedable.
 * No reference file is copied or edited; see oracle/Makefile for the recipe.
 *
 * Why each line exists:
 *  - <stack>/<tuple>: include/vptShadeMethods.h:502 uses std::stack / std::tuple without including them.
 *  - using std::abs: the reference calls unqualified abs(double) (include/Sphere.h:34,
 *    include/pathTracingUtilities.h:20 ...). With libstdc++ that binds to C int abs(int) and
 *    truncates; the author built with libc++ where it binds to the double overload.
 *  - erand48 macro (only with -DVPT_L0_TLS_RNG): the reference draws every random number from one
 *    process-global seed (include/Vector.cpp:8) shared by all OpenMP threads (a data race).  The
 *    harness redirects the calls to a thread-local, seedable, counting wrapper around the same
 *    POSIX 48-bit LCG, so results are reproducible per (seed) and threads do not interfere.
 */
#ifndef VPT_L0_SHIM_H
#define VPT_L0_SHIM_H
#ifdef __cplusplus
#include <cmath>
#include <cstdlib>
#include <stack>
#include <tuple>
#include <vector>
using std::abs;
#ifdef VPT_L0_TLS_RNG
double vpt_l0_erand48(unsigned short *ignored);
#define erand48(s) vpt_l0_erand48(s)
#endif
#endif
#endif
