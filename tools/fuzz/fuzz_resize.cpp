#include "image.h"
#include <stdio.h>
#include <random>
using namespace mg4;
int main(){ std::mt19937 rng(3); int sizes[][2]={{1,1},{1,500},{500,1},{2,3},{5,3},{223,225},{224,1},{10000,2},{3,9000},{448,448},{4000,3000}};
 for(auto&s:sizes){ std::vector<uint8_t> src((size_t)s[0]*s[1]*3); for(auto&v:src) v=rng(); std::vector<uint8_t> dst(224*224*3); resize_bicubic_u8(src.data(),s[0],s[1],dst.data(),224,224); std::vector<float> o(3*224*224); normalize_to_chw(dst.data(),224,224,o.data()); unsigned long sum=0; for(auto v:dst) sum+=v; printf("%dx%d sum %lu\n",s[0],s[1],sum);} }
