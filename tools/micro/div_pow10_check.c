// Exhaustive proof that q0 = m*r; q1 = fma(fma(-q0, p, m), r, q0) with p = 10^k, r = RN(1/p) equals the correctly
// rounded m / p for every integer 0 <= m < 10^8 and 1 <= k <= 8 (used by swar_score in csrc/ofg_parse.cuh).
// Build: gcc -O2 -fopenmp -mfma -ffp-contract=off div_pow10_check.c -lm
#include <stdio.h>
#include <math.h>
#include <stdint.h>
int main(void){
    const double p10[9]={1,1e1,1e2,1e3,1e4,1e5,1e6,1e7,1e8};
    long bad=0;
    for(int k=1;k<=8;k++){
        const double p=p10[k], r=1.0/p;
        long badk=0;
        #pragma omp parallel for reduction(+:badk)
        for(long m=0;m<100000000L;m++){
            const double a=(double)m;
            const double q0=a*r;
            const double rem=fma(-q0,p,a);
            const double q1=fma(rem,r,q0);
            if(q1!=a/p) badk++;
        }
        printf("k=%d bad=%ld\n",k,badk); bad+=badk;
    }
    printf("total bad %ld\n",bad);
    return 0;
}
