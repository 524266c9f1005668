// MiDU guidance head (SD variant) -- guidance_classifier/MiduClassifier.py:145-160.  Implemented in midu_impl below.
#include "common.cuh"
#include "rgie.h"

struct RgieMiduHead { int dummy; };

extern "C" {
int rgie_midu_create(const float* const*, int, int, int, int, int, RgieMiduHead**) { return rgie::fail("rgie_midu_create: not implemented yet"); }
void rgie_midu_destroy(RgieMiduHead*) {}
int rgie_midu_forward(RgieMiduHead*, const float*, int, float*, void*) { return rgie::fail("rgie_midu_forward: not implemented yet"); }
int rgie_midu_backward(RgieMiduHead*, const float*, float*, void*) { return rgie::fail("rgie_midu_backward: not implemented yet"); }
}
