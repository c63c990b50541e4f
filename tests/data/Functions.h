// User physics for the generated-code tests: same interface as the reference's "Unit test/Functions.h":2-4.
// Own restatement (see oracle/fv_rusanov_oracle.c for the arithmetic), dimension chosen with -DDIMENSIONS=2|3.
#pragma once
void Flux(const double* __restrict__ Q, int normal, double* __restrict__ F);
double maxEigenvalue(const double* __restrict__ Q, int normal);
double max(double* a, double* b);
