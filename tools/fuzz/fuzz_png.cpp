#include "image.h"
#include <zlib.h>
#include <stdio.h>
#include <string.h>
#include <random>
using namespace mg4;
static uint32_t be32(const uint8_t*p){return (p[0]<<24)|(p[1]<<16)|(p[2]<<8)|p[3];}
static void wbe32(uint8_t*p,uint32_t v){p[0]=v>>24;p[1]=v>>16;p[2]=v>>8;p[3]=v;}
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); std::vector<uint8_t> good(1<<20); good.resize(fread(good.data(),1,good.size(),f)); fclose(f);
  std::mt19937 rng(atoi(argv[2])); int ok=0,bad=0;
  for(int it=0;it<atoi(argv[3]);++it){
    std::vector<uint8_t> b=good;
    int nmut=1+rng()%6;
    for(int k=0;k<nmut;++k){ size_t i=8+rng()%(b.size()-8); b[i]=rng(); }
    if(it%5==0) b.resize(rng()%b.size());
    // fix chunk CRCs so the mutation reaches the decoder (walk with the possibly mutated lengths, bounded)
    size_t pos=8; while(pos+12<=b.size()){ uint32_t n=be32(&b[pos]); if(n>b.size()-pos-12) break; wbe32(&b[pos+8+n], crc32(0,&b[pos+4],n+4)); pos+=12+(size_t)n; }
    RgbImage im; std::string err;
    if(decode_png(b.data(),b.size(),im,err)) ++ok; else ++bad;
  }
  printf("ok %d bad %d\n",ok,bad); return 0; }
