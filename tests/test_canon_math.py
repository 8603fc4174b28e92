"""the canonical elementary functions (include/sac_canon_math.h) on the host: accuracy against libm in long double and
the edge cases the codec relies on. Device/host bit-equality is covered by the -m gpu parity tests (every residual
depends on them)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "sac_canon_math.h"
#include <cmath>
#include <cstdio>
#include <random>
using namespace sac_canon;
static double ulps(double a,double ref){ if(a==ref) return 0; double u=std::fabs(std::nextafter(ref,INFINITY)-ref); return std::fabs(a-ref)/u; }
int main(){
  std::mt19937_64 g(1); std::uniform_real_distribution<double> U(-40,40), V(0,1);
  double me=0,ml=0,mp=0,mp2=0,mp3=0;
  for(int i=0;i<300000;i++){
    double x=U(g); me=std::fmax(me,ulps(c_exp(x),(double)expl((long double)x)));
    double y=std::exp(U(g)); ml=std::fmax(ml,ulps(c_log(y),(double)logl((long double)y)));
    double b=0.1+V(g)*1e4, e=-(0.1+1.9*V(g)); mp=std::fmax(mp,ulps(c_pow(b,e),(double)powl(b,e)));
    double md=0.98+0.02*V(g); int n=(int)(V(g)*8192); mp2=std::fmax(mp2,ulps(c_pow(md,n),(double)powl(md,n)));
    double pd=V(g); mp3=std::fmax(mp3,ulps(c_pow(1+n,pd),(double)powl(1+n,pd)));
  }
  printf("%.3f %.3f %.3f %.3f %.3f\n",me,ml,mp,mp2,mp3);
  int bad=0;
  for(double x: {0.5,-0.5,1.5,2.5,-2.5,0.49999999999999994,1e15+0.5,-1e300,123456.5,-0.2,0.0}) if(c_round(x)!=std::round(x)) bad++;
  if(c_exp(-746)!=0.0 || c_exp(0)!=1.0 || !std::isinf(c_exp(710)) || c_pow(2.0,0.0)!=1.0 || c_pow(1.0,5.5)!=1.0) bad++;
  if(c_exp(-745.0)!=std::exp(-745.0)) bad++;
  printf("%d\n",bad);
}
'''


def test_canon_math_accuracy_and_edges():
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.cpp"); exe = os.path.join(td, "t")
        open(src, "w").write(SRC)
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe], text=True).split()
    me, ml, mp, mp2, mp3 = map(float, out[:5])
    assert me <= 1.0 and ml <= 1.0 and mp <= 2.0 and mp2 <= 4.0 and mp3 <= 2.0, out
    assert int(out[5]) == 0
