// Exhaustive host check behind dev_div_const (scvod_device_math.cuh): (float)(((double)r*180.0)/M_PI) and the deg2rad twin vs the
// fma-corrected multiplication by the rounded reciprocal, over all 2^32 float bit patterns.  g++ -O2 -mfma -pthread.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
static inline float ref_r2d(float r){ return (float)(((double)r*180.0)/M_PI); }
static const double INV_PI = 1.0/M_PI; // correctly rounded? check separately
static inline float fast_r2d(float r){ double a=(double)r*180.0; double q0=a*INV_PI; double rr=std::fma(-q0,M_PI,a); double q1=std::fma(rr,INV_PI,q0); return (float)q1; }
static inline float ref_d2r(float d){ return (float)(((double)d*M_PI)/180.0); }
static const double INV_180 = 1.0/180.0;
static inline float fast_d2r(float d){ double a=(double)d*M_PI; double q0=a*INV_180; double rr=std::fma(-q0,180.0,a); double q1=std::fma(rr,INV_180,q0); return (float)q1; }
int main(){
  int nt=8; std::vector<std::thread> th; std::atomic<long long> bad1(0), bad2(0), badd(0);
  for(int t=0;t<nt;++t) th.emplace_back([&,t]{
    long long b1=0,b2=0,bd=0;
    for(uint64_t u=t; u< (1ull<<32); u+=nt){ uint32_t v=(uint32_t)u; float r; memcpy(&r,&v,4);
      float a=ref_r2d(r), b=fast_r2d(r); uint32_t ua,ub; memcpy(&ua,&a,4); memcpy(&ub,&b,4);
      if(ua!=ub && !(a!=a && b!=b)) { ++b1; if(b1<5) printf("r2d mismatch %a: %a vs %a\n", r,a,b); }
      // double-level equality too
      double qa=((double)r*180.0)/M_PI; double aa=(double)r*180.0; double q0=aa*INV_PI; double q1=std::fma(std::fma(-q0,M_PI,aa),INV_PI,q0);
      if(qa!=q1 && !(qa!=qa)) ++bd;
      float c=ref_d2r(r), d=fast_d2r(r); memcpy(&ua,&c,4); memcpy(&ub,&d,4);
      if(ua!=ub && !(c!=c && d!=d)) { ++b2; if(b2<5) printf("d2r mismatch %a: %a vs %a\n", r,c,d); }
    }
    bad1+=b1; bad2+=b2; badd+=bd; });
  for(auto&x:th) x.join();
  printf("r2d float mismatches %lld, double-quotient mismatches %lld, d2r float mismatches %lld\n",(long long)bad1,(long long)badd,(long long)bad2);
}
