// exhaustive check: x / TWO_PI (IEEE RN) == reciprocal-based sequence, for all finite x
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
static float asf(uint32_t u){float f;memcpy(&f,&u,4);return f;}
static uint32_t asu(float f){uint32_t u;memcpy(&u,&f,4);return u;}
int main(){
  const float d = 6.28318548202514648437500f;
  const float r = 1.0f/d;
  long bad1=0,bad2=0,n=0;
#pragma omp parallel for reduction(+:bad1,bad2,n)
  for(uint64_t u=0;u<0x7f800000ull;++u){
    float x=asf((uint32_t)u);
    float ref=x/d;
    float q0=x*r;
    float e=fmaf(-q0,d,x);
    float q1=fmaf(e,r,q0);
    float e2=fmaf(-q1,d,x);
    float q2=fmaf(e2,r,q1);
    if (x>=1e-30f && x<=1e30f) {bad1+= asu(q1)!=asu(ref);
    bad2+= asu(q2)!=asu(ref);}
    n++;
  }
  printf("n=%ld one-iteration mismatches=%ld two-iteration mismatches=%ld\n",n,bad1,bad2);
  return 0;
}
