// lsd_host.h: FramePoseStruct::getCamToWorld chained along the tracking parents and the refToKf product of DepthMap.cpp:1099,
// printed for tests/test_host_adapter.py (host code only: no GPU is touched)
#include <cmath>
#include <cstdio>
#include "../../lsd_slam_b200/host/lsd_host.h"
using namespace lsd_slam;

static void setPose(Sim3& s, const double* q)
{
    double n = 0;
    for (int i = 0; i < 4; i++) n += q[i] * q[i];
    n = std::sqrt(n);
    for (int i = 0; i < 4; i++) s.q[i] = q[i] / n;
    for (int i = 0; i < 3; i++) s.t[i] = q[4 + i];
    s.s = q[7];
}
static void print(const Sim3& r)
{
    std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", r.q[0], r.q[1], r.q[2], r.q[3], r.t[0], r.t[1], r.t[2], r.s);
}

int main()
{
    // world <- keyframe 0 <- { keyframe 1, frame }: frame was tracked on keyframe 0, keyframe 1 is active
    FramePoseStruct kf0, kf1, fr;
    kf1.trackingParent = &kf0; fr.trackingParent = &kf0;
    const double a[8] = { 0.1, -0.2, 0.05, 0.97, 0.3, -0.1, 0.2, 1.7 }, b[8] = { -0.05, 0.1, 0.2, 0.97, -0.2, 0.4, 0.1, 0.9 };
    setPose(kf1.thisToParent_raw, a);
    setPose(fr.thisToParent_raw, b);
    print(kf1.getCamToWorld().inverse() * fr.getCamToWorld());
    print(kf1.thisToParent_raw);
    print(fr.thisToParent_raw);
    return 0;
}
