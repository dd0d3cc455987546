#include "../../gkr_b200/csrc/host_field.hpp"
#include <cstdio>
#include <random>
using namespace gkr;
int main(){
  std::mt19937_64 g(7);
  long bad=0;
  auto chk=[&](HFr a){ HFr s=hfr_sqr(a), m=hfr_mul(a,a); if(!hfr_eq(s,m)) ++bad; };
  HFr pm1{{hf::P[0]-1,hf::P[1],hf::P[2],hf::P[3]}};
  chk(pm1); chk(hfr_zero()); chk(hfr_one()); chk(HFr{{1,0,0,0}}); chk(HFr{{~0ull,~0ull,~0ull,0x30644e72e131a028ull}});
  chk(HFr{{~0ull,0,0,0}}); chk(HFr{{0,~0ull,0,0}}); chk(HFr{{0,0,~0ull,0}}); chk(HFr{{0,0,0,0x30644e72e131a029ull}});
  for(long i=0;i<400000;++i){ HFr a{{g(),g(),g(),g()&0x1fffffffffffffffull}}; if(hf::geq_p(a.l)) { a.l[3]>>=1; } chk(a);} 
  printf("bad=%ld\n",bad); return bad!=0; }
