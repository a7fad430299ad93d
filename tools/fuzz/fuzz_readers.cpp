#include "formats.h"
#include <stdio.h>
#include <random>
#include <fstream>
using namespace mg4;
int main(int argc,char**argv){
  const bool llama = argv[1][0]=='l';
  FILE*f=fopen(argv[2],"rb"); std::vector<unsigned char> good(200<<20); good.resize(fread(good.data(),1,good.size(),f)); fclose(f);
  std::mt19937 rng(atoi(argv[3])); int ok=0,bad=0; g_verbosity=0;
  const size_t hdr = std::min<size_t>(good.size(), 16384);
  for(int it=0;it<atoi(argv[4]);++it){
    std::vector<unsigned char> b=good;
    int mode=rng()%4;
    if(mode==0) b.resize(rng()%b.size());
    else if(mode==1){ int n=1+rng()%8; for(int k=0;k<n;++k) b[rng()%hdr]=rng(); }
    else if(mode==2){ size_t i=rng()%(hdr-4); unsigned v=rng(); memcpy(&b[i],&v,4); }
    else { size_t i=rng()%(hdr-8); unsigned long long v=((unsigned long long)rng()<<32)|rng(); if(rng()&1) v|=0x8000000000000000ull; memcpy(&b[i],&v,8); }
    FILE*o=fopen("/tmp/mg4fuzz/m.bin","wb"); fwrite(b.data(),1,b.size(),o); fclose(o);
    bool r; if(llama){ LlamaFile lf; r=lf.load("/tmp/mg4fuzz/m.bin"); } else { VisionFile vf; r=vf.load("/tmp/mg4fuzz/m.bin")==ErrNone; }
    r?++ok:++bad;
  }
  printf("ok %d bad %d\n",ok,bad); return 0; }
