// Prints results of lsd_slam_b200/csrc/hostmath.h (the product's host/device math, compiled here for the host with g++)
// for inputs read from stdin; tests/test_hostmath_cpp.py compares them with the oracle.  One line per query:
//   se3exp a0..a5 | se3mul a[7] b[7] | se3inv a[7] | sim3exp a0..a6 | sim3mul a[8] b[8] | sim3inv a[8] | sim3pose a[8]
//   ldlt6 A[36] b[6] | ldlt7 A[49] b[7] | mat3inv m[9]
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../lsd_slam_b200/csrc/hostmath.h"

template <typename T> static void out(const T* v, int n)
{
    for (int i = 0; i < n; i++) std::printf("%.17g ", (double)v[i]);
    std::printf("\n");
}
int main()
{
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream is(line);
        std::string op;
        is >> op;
        std::vector<double> v;
        double x;
        while (is >> x) v.push_back(x);
        if (op == "se3exp") {
            float a[6]; for (int i = 0; i < 6; i++) a[i] = (float)v[i];
            lsd::SE3<float> r = lsd::se3Exp(a);
            float o[7] = { r.q[0], r.q[1], r.q[2], r.q[3], r.t[0], r.t[1], r.t[2] }; out(o, 7);
        } else if (op == "se3mul" || op == "se3inv") {
            lsd::SE3<float> a, b;
            for (int i = 0; i < 4; i++) { a.q[i] = (float)v[i]; if (op == "se3mul") b.q[i] = (float)v[7 + i]; }
            for (int i = 0; i < 3; i++) { a.t[i] = (float)v[4 + i]; if (op == "se3mul") b.t[i] = (float)v[11 + i]; }
            lsd::SE3<float> r = op == "se3mul" ? lsd::se3Mul(a, b) : lsd::se3Inverse(a);
            float o[7] = { r.q[0], r.q[1], r.q[2], r.q[3], r.t[0], r.t[1], r.t[2] }; out(o, 7);
        } else if (op == "sim3exp") {
            double o[8]; lsd::sim3ToQts(lsd::sim3Exp(v.data()), o); out(o, 8);
        } else if (op == "sim3mul") {
            double o[8]; lsd::sim3ToQts(lsd::sim3Mul(lsd::sim3FromQts(v.data()), lsd::sim3FromQts(v.data() + 8)), o); out(o, 8);
        } else if (op == "sim3inv") {
            double o[8]; lsd::sim3ToQts(lsd::sim3Inverse(lsd::sim3FromQts(v.data())), o); out(o, 8);
        } else if (op == "sim3pose") {
            float R[9], t[3], roll[4]; lsd::sim3PoseConstants(lsd::sim3FromQts(v.data()), R, t, roll);
            float o[16]; for (int i = 0; i < 9; i++) o[i] = R[i]; for (int i = 0; i < 3; i++) o[9 + i] = t[i]; for (int i = 0; i < 4; i++) o[12 + i] = roll[i];
            out(o, 16);
        } else if (op == "ldlt6" || op == "ldlt7") {
            const int n = op == "ldlt6" ? 6 : 7;
            std::vector<float> A(n * n), b(n), xo(n);
            for (int i = 0; i < n * n; i++) A[i] = (float)v[i];
            for (int i = 0; i < n; i++) b[i] = (float)v[n * n + i];
            if (n == 6) lsd::ldltSolve<6>(A.data(), b.data(), xo.data()); else lsd::ldltSolve<7>(A.data(), b.data(), xo.data());
            out(xo.data(), n);
        } else if (op == "mat3inv") {
            float m[9], r[9]; for (int i = 0; i < 9; i++) m[i] = (float)v[i];
            lsd::mat3Inverse(m, r); out(r, 9);
        }
    }
    return 0;
}
