#include "image.h"
#include <stdio.h>
#include <random>
using namespace mg4;
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); std::vector<uint8_t> good(1<<20); good.resize(fread(good.data(),1,good.size(),f)); fclose(f);
  std::mt19937 rng(atoi(argv[2])); int ok=0,bad=0;
  for(int it=0;it<atoi(argv[3]);++it){
    std::vector<uint8_t> b=good;
    int nmut=1+rng()%4;
    for(int k=0;k<nmut;++k){ size_t i=2+rng()%((it&1)? std::min<size_t>(b.size()-2, 700) : b.size()-2); if(rng()%3==0) b[i]=rng(); else b[i]^=1u<<(rng()%8); }
    if(it%7==0) b.resize(2+rng()%(b.size()-2));
    RgbImage im; std::string err;
    if(decode_jpeg(b.data(),b.size(),im,err)) ++ok; else ++bad;
  }
  printf("ok %d bad %d\n",ok,bad); return 0; }
