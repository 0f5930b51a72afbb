// full-scale convergence check of the one-pass (u-state, Chronopoulos-Gear) recurrences against classic Jacobi-PCG
// on one square mixed site/bond realization (same matrix conventions as the library: every lattice bond in the matrix,
// g0 inside the spanning cluster, gleak elsewhere; rows 0 / n-1 Dirichlet).  OpenMP over rows.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>
static uint64_t sm(uint64_t *s){uint64_t z=(*s+=0x9e3779b97f4a7c15ULL);z=(z^(z>>30))*0xbf58476d1ce4e5b9ULL;z=(z^(z>>27))*0x94d049bb133111ebULL;return z^(z>>31);}
static double urand(uint64_t *s){return (sm(s)>>11)*(1.0/9007199254740992.0);}
static int *par; static int find(int a){while(par[a]!=a){par[a]=par[par[a]];a=par[a];}return a;}
int main(int argc,char**argv){
  int L=atoi(argv[1]); double ps=atof(argv[2]),pb=atof(argv[3]); uint64_t seed=atoll(argv[4]); double tol=atof(argv[5]);
  int which=argc>6?atoi(argv[6]):3;
  int m=L,n=L; int64_t t=(int64_t)m*n; double g0=1.0,gl=1e-12,Va=1.0;
  uint8_t *site=malloc(t),*bE=malloc(t),*bN=malloc(t);
  for(int64_t i=0;i<t;i++){site[i]=urand(&seed)<ps;}
  for(int64_t i=0;i<t;i++){bE[i]=urand(&seed)<pb;bN[i]=urand(&seed)<pb;}
  par=malloc(sizeof(int)*t); for(int64_t i=0;i<t;i++)par[i]=i;
  for(int y=0;y<n;y++)for(int x=0;x<m;x++){int64_t i=(int64_t)y*m+x; if(!site[i])continue;
    if(x+1<m&&bE[i]&&site[i+1]){int a=find(i),b=find(i+1);if(a!=b)par[a>b?a:b]=a>b?b:a;}
    if(y+1<n&&bN[i]&&site[i+m]){int a=find(i),b=find(i+m);if(a!=b)par[a>b?a:b]=a>b?b:a;}}
  // spanning cluster with the smallest root id
  uint8_t *bot=calloc(t,1); int cid=-1;
  for(int x=0;x<m;x++) if(site[x]) bot[find(x)]=1;
  for(int x=0;x<m;x++){int64_t i=t-m+x; if(site[i]){int r=find(i); if(bot[r]&&(cid<0||r<cid))cid=r;}}
  if(cid<0){printf("no spanning cluster\n");return 1;}
  // conducting bits: 1 E, 2 N, 4 W, 8 S
  uint8_t *cf=calloc(t,1);
  for(int y=0;y<n;y++)for(int x=0;x<m;x++){int64_t i=(int64_t)y*m+x; if(!site[i]||find(i)!=cid)continue;
    if(x+1<m&&bE[i]&&site[i+1]){cf[i]|=1;cf[i+1]|=4;}
    if(y+1<n&&bN[i]&&site[i+m]){cf[i]|=2;cf[i+m]|=8;}}
  free(par);free(bot);free(site);free(bE);free(bN);
  double *d=malloc(8*t),*inv=malloc(8*t),*b=calloc(t,8);
  #pragma omp parallel for
  for(int y=0;y<n;y++)for(int x=0;x<m;x++){int64_t i=(int64_t)y*m+x; int ne=(x+1<m)+(x>0)+(y+1<n)+(y>0); int nc=__builtin_popcount(cf[i]);
    d[i]=fma((double)(ne-nc),gl,(double)nc*g0); inv[i]=1.0/d[i];}
  double bn=0; for(int x=0;x<m;x++){int64_t i=(int64_t)(n-2)*m+x; b[i]=((cf[i]&2)?g0:gl)*Va; double z=b[i]*inv[i]; bn+=z*z;} bn=sqrt(bn);
  #define WGT(i,bit) ((cf[i]&(bit))?g0:gl)
  // ---------------- classic Jacobi-PCG
  if(which&1){
  double *x_=calloc(t,8),*r=malloc(8*t),*p=calloc(t,8),*q=calloc(t,8); memcpy(r,b,8*t);
  double bknum=0; for(int64_t i=m;i<t-m;i++)bknum+=r[i]*r[i]*inv[i];
  double bk=0,err=0; int it=0; double t0=omp_get_wtime();
  for(;;){
    #pragma omp parallel for
    for(int64_t i=m;i<t-m;i++)p[i]=r[i]*inv[i]+bk*p[i];
    double den=0;
    #pragma omp parallel for reduction(+:den)
    for(int y=1;y<n-1;y++)for(int x=0;x<m;x++){int64_t i=(int64_t)y*m+x; double a=d[i]*p[i];
      if(x+1<m)a-=WGT(i,1)*p[i+1]; if(x>0)a-=WGT(i,4)*p[i-1]; a-=WGT(i,2)*p[i+m]; a-=WGT(i,8)*p[i-m]; q[i]=a; den+=p[i]*a;}
    double ak=bknum/den, rz=0,rr=0;
    #pragma omp parallel for reduction(+:rz,rr)
    for(int64_t i=m;i<t-m;i++){x_[i]+=ak*p[i]; double v=r[i]-ak*q[i]; r[i]=v; rz+=v*v*inv[i]; rr+=v*v;}
    it++; err=sqrt(rr)/bn; bk=rz/bknum; bknum=rz;
    if(it%5000==0){printf("  classic it=%d err=%.3e (%.1fs)\n",it,err,omp_get_wtime()-t0);fflush(stdout);}
    if(!(err>tol)||it>4000000)break;}
  double Itop=0; for(int x=0;x<m;x++){int64_t i=t-m+x; double a=d[i]*Va; if(x+1<m&&WGT(i,1)>=1e-10)a-=WGT(i,1)*Va; if(x>0&&WGT(i,4)>=1e-10)a-=WGT(i,4)*Va; if(WGT(i,8)>=1e-10)a-=WGT(i,8)*x_[i-m]; Itop+=a;}
  printf("classic : iters=%d err=%.3e Gtop=%.12e  (%.1f s)\n",it,err,Itop/Va,omp_get_wtime()-t0);fflush(stdout);
  free(x_);free(r);free(p);free(q);}
  // ---------------- one-pass recurrences, state u = r/d and s (as pcg_fused_kernel, variant FtCfgA3)
  if(which&2){
  double *u=calloc(t,8),*u2=calloc(t,8),*s=calloc(t,8),*s2=calloc(t,8),*xr=calloc(m,8),*pr=calloc(m,8);
  for(int64_t i=m;i<t-m;i++)u[i]=b[i]*inv[i];
  double alpha=0,beta=0,gamma=0; int it=0,prime=1; double err=0; double t0=omp_get_wtime();
  for(;;){
    double rz=0,rr=0;
    #pragma omp parallel for reduction(+:rz,rr)
    for(int y=1;y<n-1;y++)for(int x=0;x<m;x++){int64_t i=(int64_t)y*m+x; double a=d[i]*u[i];
      double all=0,con=0; 
      if(x+1<m){all+=u[i+1]; if(cf[i]&1)con+=u[i+1];} if(x>0){all+=u[i-1]; if(cf[i]&4)con+=u[i-1];}
      all+=u[i+m]; if(cf[i]&2)con+=u[i+m]; all+=u[i-m]; if(cf[i]&8)con+=u[i-m];
      double w=a-(gl*all+(g0-gl)*con);
      double sn=w+beta*s[i]; double un=u[i]-alpha*(sn*inv[i]); double rn=d[i]*un;
      s2[i]=sn; u2[i]=un; rz+=rn*un; rr+=rn*rn;
      if(y==n-2){double pp=u[i]+beta*pr[x]; pr[x]=pp; xr[x]+=alpha*pp;}}
    // delta' = energy of the bonds (u2 is zero on the Dirichlet rows)
    double en=0;
    #pragma omp parallel for reduction(+:en)
    for(int y=0;y<n-1;y++)for(int x=0;x<m;x++){int64_t i=(int64_t)y*m+x; double all=0,con=0;
      if(y>=1&&x+1<m){double e=u2[i]-u2[i+1]; e*=e; all+=e; if(cf[i]&1)con+=e;}
      {double e=u2[i]-u2[i+m]; e*=e; all+=e; if(cf[i]&2)con+=e;}
      en+=gl*all+(g0-gl)*con;}
    double *tmp=u;u=u2;u2=tmp; tmp=s;s=s2;s2=tmp;
    if(prime){gamma=rz; alpha=rz/en; beta=0; prime=0; continue;}
    it++; err=sqrt(rr)/bn;
    if(it%5000==0){printf("  one-pass it=%d err=%.3e (%.1fs)\n",it,err,omp_get_wtime()-t0);fflush(stdout);}
    if(!(err>tol)||it>4000000)break;
    beta=rz/gamma; alpha=rz/(en-beta*rz/alpha); gamma=rz;}
  double Itop=0; for(int x=0;x<m;x++){int64_t i=t-m+x; double a=d[i]*Va; if(x+1<m&&WGT(i,1)>=1e-10)a-=WGT(i,1)*Va; if(x>0&&WGT(i,4)>=1e-10)a-=WGT(i,4)*Va; if(WGT(i,8)>=1e-10)a-=WGT(i,8)*xr[x]; Itop+=a;}
  printf("one-pass: iters=%d err=%.3e Gtop=%.12e  (%.1f s)\n",it,err,Itop/Va,omp_get_wtime()-t0);fflush(stdout);}
  return 0;}
