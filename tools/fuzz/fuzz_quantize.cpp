#include "formats.h"
#include <stdio.h>
#include <string.h>
#include <random>
extern "C" int mg4_quantize_container(const char *in_path, const char *out_path, int data_type);
using namespace mg4;
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); std::vector<unsigned char> good(200<<20); good.resize(fread(good.data(),1,good.size(),f)); fclose(f);
  std::mt19937 rng(atoi(argv[2])); int codes[32]={0}; g_verbosity=0;
  const size_t hdr = std::min<size_t>(good.size(), 16384);
  for(int it=0;it<atoi(argv[3]);++it){
    std::vector<unsigned char> b=good;
    int mode=rng()%4;
    if(mode==0) b.resize(rng()%b.size());
    else if(mode==1){ int n=1+rng()%4; for(int k=0;k<n;++k) b[rng()%hdr]=rng(); }
    else if(mode==2){ size_t i=rng()%(hdr-4); unsigned v=rng(); memcpy(&b[i],&v,4); }
    FILE*o=fopen("/tmp/mg4fuzz/qm.bin","wb"); fwrite(b.data(),1,b.size(),o); fclose(o);
    int types[5]={4,5,6,7,8};
    int rc=mg4_quantize_container("/tmp/mg4fuzz/qm.bin","/tmp/mg4fuzz/qo.bin",types[rng()%5]);
    codes[rc&31]++;
  }
  for(int i=0;i<32;++i) if(codes[i]) printf("code %d: %d\n",i,codes[i]); return 0; }
