// tools/quantiser_check.c — exhaustive proof-by-enumeration that the integer quantiser used by gc_encode.cu equals the
// reference cast chain (GcAdpcmEncoder.cs:142-144) for every scale shift K=11..23: all |d| <= 2^25, every rounding
// threshold of the float32 conversion above 2^24, 2e8 random int32 per K and the int32 extremes.
// Build/run: gcc -O2 -ffp-contract=off -o /tmp/qc tools/quantiser_check.c && /tmp/qc   (~20 s, exit code 0 = identical)
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
static inline int32_t lit(int32_t diff, int K){ float scale=(float)(1u<<K); float ratio=(float)diff/scale;
  return diff>0? (int32_t)((double)ratio+(double)0.4999999f):(int32_t)((double)ratio-(double)0.4999999f);}
// exact integer form: round half toward zero of RNE24(|d|)/2^K
static inline int32_t fastq(int32_t d, int K){ uint32_t a = d<0? (uint32_t)(-(int64_t)d):(uint32_t)d; uint32_t t=a>>24; uint32_t hs = t? (1u<<(31-__builtin_clz(t))):0;
  uint32_t m = (a + (1u<<(K-1)) - 1 - hs) >> K; return d<0? -(int32_t)m:(int32_t)m; }
// signed small form valid for |d|<2^24
static inline int32_t smallq(int32_t d,int K){ return (d + (1<<(K-1)) - (d>0)) >> K; }
int main(){ long bad=0;
  for(int K=11;K<=23;K++){
    for(int32_t d=-(1<<25); d<=(1<<25); d++){ int32_t l=lit(d,K); if(fastq(d,K)!=l){ if(bad++<10)printf("fast mismatch K=%d d=%d lit=%d fast=%d\n",K,d,l,fastq(d,K));}
      if(d>-(1<<24)&&d<(1<<24)&&smallq(d,K)!=l){ if(bad++<10)printf("small mismatch K=%d d=%d lit=%d small=%d\n",K,d,l,smallq(d,K));} }
    // thresholds for large d
    for(int64_t j=1;j<=(1ll<<(32-K));j++){ int64_t T=j*(1ll<<K)-(1ll<<(K-1)); for(int s=0;s<=8;s++) for(int dd=-3;dd<=3;dd++){ int64_t v=T+(s?(1ll<<(s-1)):0)+dd; if(v<=0||v>2147483647ll) continue; int32_t d=(int32_t)v; if(fastq(d,K)!=lit(d,K)){if(bad++<10)printf("thr mismatch K=%d d=%d\n",K,d);} d=-d; if(fastq(d,K)!=lit(d,K)){if(bad++<10)printf("thr mismatch K=%d d=%d\n",K,d);} } }
    uint64_t x=88172645463325252ull; for(long i=0;i<200000000;i++){ x^=x<<13;x^=x>>7;x^=x<<17; int32_t d=(int32_t)x; if(fastq(d,K)!=lit(d,K)){if(bad++<10)printf("rand mismatch K=%d d=%d lit=%d fast=%d\n",K,d,lit(d,K),fastq(d,K));} }
    int32_t ed[]={INT32_MIN,INT32_MIN+1,INT32_MAX,INT32_MAX-1,INT32_MAX-63,INT32_MAX-64,INT32_MAX-65,0,1,-1};
    for(unsigned i=0;i<sizeof ed/4;i++) if(fastq(ed[i],K)!=lit(ed[i],K)){bad++;printf("edge mismatch K=%d d=%d lit=%d fast=%d\n",K,ed[i],lit(ed[i],K),fastq(ed[i],K));}
    printf("K=%d done bad=%ld\n",K,bad); fflush(stdout);
  } return bad!=0; }
