// QUOTED FROM THE REFERENCE, not original work: Rand_Uniform / Rand_Norm below restate YongonY/NBLDPC Rand.cpp:17-38 almost
// token for token, because parity needs the harness to feed the decoder bit-identical inputs -- the constants and the order
// of the floating-point operations cannot differ: three multiplicative congruential generators (multipliers 249 / 251 / 252,
// moduli 61967 / 63443 / 63599) summed modulo 1, and the cosine branch of Box-Muller with u1 drawn before u2.
// Used by the parity harness only; the device Monte-Carlo draws its noise with Philox (csrc/decoder.cu).
#pragma once
#include <cmath>

class CRand {
public:
    unsigned long IX = 0, IY = 0, IZ = 0;
    double Rand_Uniform(void) {
        IX = (IX * 249) % 61967;
        IY = (IY * 251) % 63443;
        IZ = (IZ * 252) % 63599;
        double t = ((double)IX) / ((double)61967) + ((double)IY) / ((double)63443) + ((double)IZ) / ((double)63599);
        t -= (int)t;
        return t;
    }
    double Rand_Norm(double mu, double sigma) {
        const double u1 = Rand_Uniform();
        const double u2 = Rand_Uniform();
        return mu + sigma * cos(2 * acos(-1.0) * u2) * sqrt(-2.0 * log(1.0 - u1));
    }
};
