/*
 * go1_physics_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * fp64, one-env-at-a-time restatement of the rigid-body step that the CUDA kernel
 * (walk-these-ways_b200/csrc/sim_step.cu) runs in fp32 with 4 lanes per env.
 *
 * PARITY UNPINNED: the reference delegates this step to NVIDIA Isaac Gym / PhysX
 * (legged_robot.py:76-80 `gym.simulate`), a closed binary that is not in /root/reference and not
 * installable.  No file of the reference pins its numerical results.  This oracle
 * therefore restates OUR published algorithm: Featherstone articulated-body
 * algorithm for a floating base + 12 revolute joints (model constants compiled from the reference's
 * go1.urdf by tools/compile_model.py), semi-implicit Euler at dt = 5 ms (legged_robot_config.py:402),
 * velocity-level frictional contact for the 4 foot spheres solved by projected block-Jacobi/Gauss-Seidel
 * on the 12x12 Delassus matrix, explicit penalty contacts for trunk corners / hips / knees / calves,
 * implicit spring-damper joint limits.  It is validated by invariants (tests/test_physics_oracle.py:
 * ABA vs independent RNEA, momentum/energy conservation, static equilibrium) — not against PhysX.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link this file.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>
#include <unistd.h>
#include "../walk-these-ways_b200/csrc/go1_model_generated.h"
#include "go1_physics_oracle.h"

/* ---------- small linear algebra ---------- */
static void v3cross(const double a[3], const double b[3], double o[3]) {
    double x = a[1]*b[2]-a[2]*b[1], y = a[2]*b[0]-a[0]*b[2], z = a[0]*b[1]-a[1]*b[0];
    o[0]=x; o[1]=y; o[2]=z;
}
static double v3dot(const double a[3], const double b[3]) { return a[0]*b[0]+a[1]*b[1]+a[2]*b[2]; }
static void m3v(const double M[9], const double v[3], double o[3]) {
    double t[3];
    for (int i=0;i<3;i++) t[i]=M[3*i]*v[0]+M[3*i+1]*v[1]+M[3*i+2]*v[2];
    o[0]=t[0];o[1]=t[1];o[2]=t[2];
}
static void m3tv(const double M[9], const double v[3], double o[3]) {
    double t[3];
    for (int i=0;i<3;i++) t[i]=M[i]*v[0]+M[3+i]*v[1]+M[6+i]*v[2];
    o[0]=t[0];o[1]=t[1];o[2]=t[2];
}
static void m3m(const double A[9], const double B[9], double O[9]) {
    double T[9];
    for (int i=0;i<3;i++) for (int j=0;j<3;j++) T[3*i+j]=A[3*i]*B[j]+A[3*i+1]*B[3+j]+A[3*i+2]*B[6+j];
    memcpy(O,T,sizeof T);
}
static void quat_to_R(const double q[4], double R[9]) { /* xyzw, body->world */
    double x=q[0],y=q[1],z=q[2],w=q[3];
    R[0]=1-2*(y*y+z*z); R[1]=2*(x*y-z*w);   R[2]=2*(x*z+y*w);
    R[3]=2*(x*y+z*w);   R[4]=1-2*(x*x+z*z); R[5]=2*(y*z-x*w);
    R[6]=2*(x*z-y*w);   R[7]=2*(y*z+x*w);   R[8]=1-2*(x*x+y*y);
}
static void axis_R(int axis, double q, double R[9]) { /* child->parent rotation by q about axis */
    double c=cos(q), s=sin(q);
    for (int i=0;i<9;i++) R[i]=0;
    if (axis==0) { R[0]=1; R[4]=c; R[5]=-s; R[7]=s; R[8]=c; }
    else if (axis==1) { R[0]=c; R[2]=s; R[4]=1; R[6]=-s; R[8]=c; }
    else { R[0]=c; R[1]=-s; R[3]=s; R[4]=c; R[8]=1; }
}
/* spatial vectors are [angular(3); linear(3)] in body coordinates */
static void crm(const double v[6], const double m[6], double o[6]) { /* v x m (motion) */
    double a[3], b[3], c[3];
    v3cross(v, m, a); v3cross(v, m+3, b); v3cross(v+3, m, c);
    o[0]=a[0];o[1]=a[1];o[2]=a[2]; o[3]=b[0]+c[0];o[4]=b[1]+c[1];o[5]=b[2]+c[2];
}
static void crf(const double v[6], const double f[6], double o[6]) { /* v x* f (force) */
    double a[3], b[3], c[3];
    v3cross(v, f, a); v3cross(v+3, f+3, b); v3cross(v, f+3, c);
    o[0]=a[0]+b[0];o[1]=a[1]+b[1];o[2]=a[2]+b[2]; o[3]=c[0];o[4]=c[1];o[5]=c[2];
}
static void mat6v(const double M[36], const double v[6], double o[6]) {
    double t[6];
    for (int i=0;i<6;i++){ t[i]=0; for(int j=0;j<6;j++) t[i]+=M[6*i+j]*v[j]; }
    memcpy(o,t,sizeof t);
}
/* spatial inertia about body origin from mass, com, inertia about com */
static void spatial_inertia(double m, const double c[3], const double Ic[9], double I[36]) {
    double cx[9]={0,-c[2],c[1], c[2],0,-c[0], -c[1],c[0],0};
    double cc=v3dot(c,c);
    for (int i=0;i<36;i++) I[i]=0;
    for (int i=0;i<3;i++) for(int j=0;j<3;j++){
        I[6*i+j]=Ic[3*i+j]+m*((i==j?cc:0)-c[i]*c[j]);
        I[6*i+3+j]=m*cx[3*i+j];
        I[6*(3+i)+j]=m*cx[3*j+i];
        I[6*(3+i)+3+j]=(i==j)?m:0;
    }
}
/* motion transform parent->child: E (parent coords -> child coords), r = child origin in parent */
static void xmotion(const double E[9], const double r[3], const double vp[6], double vc[6]) {
    double w[3], t[3], l[3];
    m3v(E, vp, w);
    v3cross(vp, r, t);               /* w_p x r */
    for (int i=0;i<3;i++) t[i]+=vp[3+i];
    m3v(E, t, l);
    for (int i=0;i<3;i++){ vc[i]=w[i]; vc[3+i]=l[i]; }
}
/* force transform child->parent (X^T) */
static void xforce_t(const double E[9], const double r[3], const double fc[6], double fp[6]) {
    double n[3], f[3], t[3];
    m3tv(E, fc, n); m3tv(E, fc+3, f); v3cross(r, f, t);
    for (int i=0;i<3;i++){ fp[i]=n[i]+t[i]; fp[3+i]=f[i]; }
}
static void xmat(const double E[9], const double r[3], double X[36]) { /* 6x6 motion transform */
    double rx[9]={0,-r[2],r[1], r[2],0,-r[0], -r[1],r[0],0}, Erx[9];
    m3m(E, rx, Erx);
    for (int i=0;i<36;i++) X[i]=0;
    for (int i=0;i<3;i++) for (int j=0;j<3;j++){ X[6*i+j]=E[3*i+j]; X[6*(3+i)+3+j]=E[3*i+j]; X[6*(3+i)+j]=-Erx[3*i+j]; }
}
static void xtIx_add(const double X[36], const double I[36], double O[36]) { /* O += X^T I X */
    double T[36];
    for (int i=0;i<6;i++) for(int j=0;j<6;j++){ double s=0; for(int k=0;k<6;k++) s+=I[6*i+k]*X[6*k+j]; T[6*i+j]=s; }
    for (int i=0;i<6;i++) for(int j=0;j<6;j++){ double s=0; for(int k=0;k<6;k++) s+=X[6*k+i]*T[6*k+j]; O[6*i+j]+=s; }
}
static void solve6(const double A[36], const double b[6], double x[6]) { /* SPD solve via LDL^T */
    double L[36]={0}, D[6], y[6];
    for (int j=0;j<6;j++){
        double d=A[6*j+j]; for(int k=0;k<j;k++) d-=L[6*j+k]*L[6*j+k]*D[k];
        D[j]=d; L[6*j+j]=1;
        for (int i=j+1;i<6;i++){ double s=A[6*i+j]; for(int k=0;k<j;k++) s-=L[6*i+k]*L[6*j+k]*D[k]; L[6*i+j]=s/d; }
    }
    for (int i=0;i<6;i++){ double s=b[i]; for(int k=0;k<i;k++) s-=L[6*i+k]*y[k]; y[i]=s; }
    for (int i=0;i<6;i++) y[i]/=D[i];
    for (int i=5;i>=0;i--){ double s=y[i]; for(int k=i+1;k<6;k++) s-=L[6*k+i]*x[k]; x[i]=s; }
}

/* ---------- model ---------- */
typedef struct { double r[3]; int axis; double lo, hi, vmax; double I[36]; } Joint;
static void leg_joint(int L, int j, Joint* J) {
    const double (*org)[3] = j==0?GO1_HIP_ORIGIN:(j==1?GO1_THIGH_ORIGIN:GO1_CALF_ORIGIN);
    const double* mass = j==0?GO1_HIP_MASS:(j==1?GO1_THIGH_MASS:GO1_CALF_MASS);
    const double (*com)[3] = j==0?GO1_HIP_COM:(j==1?GO1_THIGH_COM:GO1_CALF_COM);
    const double (*Ic)[9] = j==0?GO1_HIP_INERTIA_COM:(j==1?GO1_THIGH_INERTIA_COM:GO1_CALF_INERTIA_COM);
    const double (*lim)[2] = j==0?GO1_HIP_LIMITS:(j==1?GO1_THIGH_LIMITS:GO1_CALF_LIMITS);
    memcpy(J->r, org[L], sizeof J->r);
    J->axis = j==0?GO1_HIP_AXIS:(j==1?GO1_THIGH_AXIS:GO1_CALF_AXIS);
    J->lo = lim[L][0]; J->hi = lim[L][1];
    J->vmax = j==0?GO1_HIP_VEL_LIMIT:(j==1?GO1_THIGH_VEL_LIMIT:GO1_CALF_VEL_LIMIT);
    spatial_inertia(mass[L], com[L], Ic[L], J->I);
}

void go1_oracle_default_params(Go1PhysParams* P) {
    memset(P, 0, sizeof *P);
    P->dt = 0.005;
    P->gravity[2] = -9.8;                /* legged_robot.py:558 */
    P->erp = 0.2; P->cfm = 1e-4; P->max_depen_vel = 1.0;   /* legged_robot_config.py:416 */
    P->contact_margin = 0.01;            /* contact_offset, legged_robot_config.py:413 */
    P->bounce_threshold = 0.5;           /* legged_robot_config.py:415 */
    P->pgs_iters = 8;
    P->terrain_friction = 1.0; P->terrain_restitution = 0.0;   /* legged_robot_config.py:68-70 */
    P->pen_k[0]=20000; P->pen_c[0]=150;  /* base corners */
    P->pen_k[1]=20000; P->pen_c[1]=150;  /* hips */
    P->pen_k[2]=5000;  P->pen_c[2]=30;   /* knees (thigh) */
    P->pen_k[3]=5000;  P->pen_c[3]=30;   /* calf mid */
    P->pen_mt = 0.2;
    P->limit_k = 300; P->limit_c = 3;
}

/* terrain height + normal (flat when hf == NULL) */
static double terrain(const Go1PhysParams* P, double x, double y, double n[3]) {
    n[0]=0;n[1]=0;n[2]=1;
    if (!P->hf) return 0.0;
    double fx=(x+P->hf_border)/P->hf_hscale, fy=(y+P->hf_border)/P->hf_hscale;
    if (fx<0) fx=0;
    if (fy<0) fy=0;
    if (fx>P->hf_rows-1.001) fx=P->hf_rows-1.001;
    if (fy>P->hf_cols-1.001) fy=P->hf_cols-1.001;
    int ix=(int)fx, iy=(int)fy; double ax=fx-ix, ay=fy-iy;
    double h00=P->hf[ix*P->hf_cols+iy], h10=P->hf[(ix+1)*P->hf_cols+iy];
    double h01=P->hf[ix*P->hf_cols+iy+1], h11=P->hf[(ix+1)*P->hf_cols+iy+1];
    double h=(h00*(1-ax)*(1-ay)+h10*ax*(1-ay)+h01*(1-ax)*ay+h11*ax*ay)*P->hf_vscale;
    double dhdx=((h10-h00)*(1-ay)+(h11-h01)*ay)*P->hf_vscale/P->hf_hscale;
    double dhdy=((h01-h00)*(1-ax)+(h11-h10)*ax)*P->hf_vscale/P->hf_hscale;
    double inv=1.0/sqrt(dhdx*dhdx+dhdy*dhdy+1.0);
    n[0]=-dhdx*inv; n[1]=-dhdy*inv; n[2]=inv;
    return h;
}

typedef struct {
    double E[12][9], Rw[13][9], pw[13][3];   /* joint E (parent->child coords); body->world rot; origin */
    double v[13][6], c[13][6], pA[13][6], IA[13][36], U[12][6], D[12], u[12];
    Joint J[12];
} Work;

static int body_of(int L, int j) { return 1+3*L+j; }
static int parent_of(int L, int j) { return j==0?0:body_of(L,j-1); }

/* impulse/acceleration propagation (pass 2 + base solve + pass 3) for given bias forces.
   pA_in[13][6]: bias force per body; tau_u: if non-NULL, use u = tau_u[i] - S^T pA (with c terms), else pure
   impulse response (c = 0, u = -S^T pA).  Outputs base accel a0[6] and qdd[12]. */
static void propagate(const Work* W, double pA[13][6], const double* tau_eff, int with_c,
                      double a0[6], double qdd[12], double a_out[13][6]) {
    double u[12];
    for (int L=0;L<4;L++) for (int j=2;j>=0;j--) {
        int i=3*L+j, b=body_of(L,j), p=parent_of(L,j);
        const Joint* J=&W->J[i];
        u[i]=(tau_eff?tau_eff[i]:0.0)-pA[b][J->axis];
        double pa[6];
        for (int k=0;k<6;k++) pa[k]=pA[b][k]+W->U[i][k]*u[i]/W->D[i];
        if (with_c) { /* pa += Ia c ;  Ia = IA - U U^T/D */
            double Ic[6]; mat6v(W->IA[b], W->c[b], Ic);
            double Uc=0; for(int k=0;k<6;k++) Uc+=W->U[i][k]*W->c[b][k];
            for (int k=0;k<6;k++) pa[k]+=Ic[k]-W->U[i][k]*Uc/W->D[i];
        }
        double fp[6]; xforce_t(W->E[i], J->r, pa, fp);
        for (int k=0;k<6;k++) pA[p][k]+=fp[k];
    }
    double nb[6]; for(int k=0;k<6;k++) nb[k]=-pA[0][k];
    solve6(W->IA[0], nb, a0);
    double a[13][6]; memcpy(a[0], a0, sizeof a[0]);
    for (int L=0;L<4;L++) for (int j=0;j<3;j++) {
        int i=3*L+j, b=body_of(L,j), p=parent_of(L,j);
        const Joint* J=&W->J[i];
        xmotion(W->E[i], J->r, a[p], a[b]);
        if (with_c) for(int k=0;k<6;k++) a[b][k]+=W->c[b][k];
        double Ua=0; for(int k=0;k<6;k++) Ua+=W->U[i][k]*a[b][k];
        qdd[i]=(u[i]-Ua)/W->D[i];
        a[b][J->axis]+=qdd[i];
    }
    if (a_out) memcpy(a_out, a, sizeof a);
}

/* world linear velocity of a point pt (body coords) on body b given spatial velocities v[][] */
static void point_vel_world(const Work* W, const double v[13][6], int b, const double pt[3], double out[3]) {
    double t[3]; v3cross(v[b], pt, t);
    for (int k=0;k<3;k++) t[k]+=v[b][3+k];
    m3v(W->Rw[b], t, out);
}
/* spatial velocities of all bodies from base twist (body coords) + joint rates */
static void velocities(const Work* W, const double v0[6], const double qd[12], double v[13][6]) {
    memcpy(v[0], v0, sizeof v[0]);
    for (int L=0;L<4;L++) for (int j=0;j<3;j++) {
        int i=3*L+j, b=body_of(L,j), p=parent_of(L,j);
        xmotion(W->E[i], W->J[i].r, v[p], v[b]);
        v[b][W->J[i].axis]+=qd[i];
    }
}

void go1_oracle_substep(const Go1PhysParams* P, const Go1PhysDR* dr, Go1PhysState* s,
                        const double tau[12], Go1PhysOut* out) {
    Work* W=(Work*)calloc(1,sizeof(Work));
    const double dt=P->dt;
    double R0[9]; quat_to_R(s->quat, R0);
    memcpy(W->Rw[0], R0, sizeof R0); memcpy(W->pw[0], s->pos, sizeof s->pos);
    double v0[6]; m3tv(R0, s->angvel, v0); m3tv(R0, s->linvel, v0+3);

    /* base inertia: mass = default + payload, com = com displacement (legged_robot.py:667-673),
       rotational inertia scaled with the mass ratio (our reading of recomputeInertia=True) */
    double mb=GO1_BASE_MASS+dr->payload, Icb[9];
    for (int k=0;k<9;k++) Icb[k]=GO1_BASE_INERTIA_COM[k]*(mb/GO1_BASE_MASS);
    spatial_inertia(mb, dr->com_disp, Icb, W->IA[0]);
    double I0[36]; memcpy(I0, W->IA[0], sizeof I0);

    /* ---- pass 1: kinematics, velocities, bias forces ---- */
    memcpy(W->v[0], v0, sizeof v0);
    for (int L=0;L<4;L++) for (int j=0;j<3;j++) {
        int i=3*L+j, b=body_of(L,j), p=parent_of(L,j);
        leg_joint(L,j,&W->J[i]);
        double Rj[9]; axis_R(W->J[i].axis, s->q[i], Rj);
        for (int a=0;a<3;a++) for(int c2=0;c2<3;c2++) W->E[i][3*a+c2]=Rj[3*c2+a];
        m3m(W->Rw[p], Rj, W->Rw[b]);
        double t[3]; m3v(W->Rw[p], W->J[i].r, t);
        for (int k=0;k<3;k++) W->pw[b][k]=W->pw[p][k]+t[k];
        xmotion(W->E[i], W->J[i].r, W->v[p], W->v[b]);
        double Sq[6]={0,0,0,0,0,0}; Sq[W->J[i].axis]=s->qd[i];
        W->v[b][W->J[i].axis]+=s->qd[i];
        crm(W->v[b], Sq, W->c[b]);
        memcpy(W->IA[b], W->J[i].I, sizeof W->J[i].I);
    }
    double Iv[6];
    mat6v(I0, W->v[0], Iv); crf(W->v[0], Iv, W->pA[0]);
    for (int b=1;b<13;b++){ mat6v(W->IA[b], W->v[b], Iv); crf(W->v[b], Iv, W->pA[b]); }

    memset(out->contact_force, 0, sizeof out->contact_force);
    const double mu=0.5*(dr->friction+P->terrain_friction);          /* PhysX default combine: average */
    const double rest=0.5*(dr->restitution+P->terrain_restitution);

    /* ---- explicit penalty contacts: 2 trunk corners + hip + knee + calf-mid per leg ---- */
    for (int L=0;L<4;L++) {
        double sx=(L<2)?1.0:-1.0, sy=(L%2==0)?1.0:-1.0;
        struct { int body, isaac_body, cls; double pt[3], rad; } pts[5] = {
            {0, 0, 0, {sx*GO1_BASE_BOX_HALF[0], sy*GO1_BASE_BOX_HALF[1],  GO1_BASE_BOX_HALF[2]}, 0.0},
            {0, 0, 0, {sx*GO1_BASE_BOX_HALF[0], sy*GO1_BASE_BOX_HALF[1], -GO1_BASE_BOX_HALF[2]}, 0.0},
            {body_of(L,0), 1+4*L, 1, {GO1_HIP_COLL_OFFSET[L][0], GO1_HIP_COLL_OFFSET[L][1], GO1_HIP_COLL_OFFSET[L][2]}, GO1_HIP_COLL_RADIUS},
            {body_of(L,1), 2+4*L, 2, {GO1_CALF_ORIGIN[L][0], GO1_CALF_ORIGIN[L][1], GO1_CALF_ORIGIN[L][2]}, 0.017},
            {body_of(L,2), 3+4*L, 3, {0.5*GO1_FOOT_OFFSET[L][0], 0.5*GO1_FOOT_OFFSET[L][1], 0.5*GO1_FOOT_OFFSET[L][2]}, 0.008},
        };
        for (int k=0;k<5;k++) {
            int b=pts[k].body; double pwp[3], t[3], n[3];
            m3v(W->Rw[b], pts[k].pt, t); for(int a=0;a<3;a++) pwp[a]=W->pw[b][a]+t[a];
            double h=terrain(P, pwp[0], pwp[1], n);
            double gap=(pwp[2]-h)*n[2]-pts[k].rad;
            if (gap>=0) continue;
            double vw[3]; point_vel_world(W, W->v, b, pts[k].pt, vw);
            double vn=v3dot(vw,n);
            double fn=P->pen_k[pts[k].cls]*(-gap)-P->pen_c[pts[k].cls]*vn; if (fn<0) fn=0;
            double vt[3]; for(int a=0;a<3;a++) vt[a]=vw[a]-vn*n[a];
            double vtn=sqrt(v3dot(vt,vt));
            double ct=0; if (vtn>1e-9){ ct=mu*fn/vtn; double cmax=P->pen_mt/dt; if (ct>cmax) ct=cmax; }
            double F[3]; for(int a=0;a<3;a++) F[a]=fn*n[a]-ct*vt[a];
            for (int a=0;a<3;a++) out->contact_force[pts[k].isaac_body][a]+=F[a];
            double fb[3], nb[3]; m3tv(W->Rw[b], F, fb); v3cross(pts[k].pt, fb, nb);
            for (int a=0;a<3;a++){ W->pA[b][a]-=nb[a]; W->pA[b][3+a]-=fb[a]; }
        }
    }

    /* ---- pass 2 quantities that do not depend on forces: IA, U, D (with implicit joint-limit terms) ---- */
    double tau_eff[12];
    for (int L=0;L<4;L++) for (int j=2;j>=0;j--) {
        int i=3*L+j, b=body_of(L,j), p=parent_of(L,j);
        const Joint* J=&W->J[i];
        double arm=0; tau_eff[i]=tau[i];
        double viol=0; if (s->q[i]>J->hi) viol=s->q[i]-J->hi; else if (s->q[i]<J->lo) viol=s->q[i]-J->lo;
        if (viol!=0) {
            arm=dt*P->limit_c+dt*dt*P->limit_k;
            tau_eff[i]-=P->limit_c*s->qd[i]+P->limit_k*(viol+dt*s->qd[i]);
        }
        for (int k=0;k<6;k++) W->U[i][k]=W->IA[b][6*k+J->axis];
        W->D[i]=W->U[i][J->axis]+arm;
        double Ia[36];
        for (int a=0;a<6;a++) for(int c2=0;c2<6;c2++) Ia[6*a+c2]=W->IA[b][6*a+c2]-W->U[i][a]*W->U[i][c2]/W->D[i];
        double X[36]; xmat(W->E[i], J->r, X);
        xtIx_add(X, Ia, W->IA[p]);
    }

    /* ---- free (contact-free) accelerations ---- */
    double pA[13][6]; memcpy(pA, W->pA, sizeof pA);
    double a0[6], qdd[12];
    propagate(W, pA, tau_eff, 1, a0, qdd, NULL);
    double gb[3]; m3tv(R0, P->gravity, gb);
    double wxv[3]; v3cross(v0, v0+3, wxv);
    double vf0[6], qdf[12];
    for (int k=0;k<3;k++){ vf0[k]=v0[k]+dt*a0[k]; vf0[3+k]=v0[3+k]+dt*(a0[3+k]+gb[k]+wxv[k]); }
    for (int i=0;i<12;i++) qdf[i]=s->qd[i]+dt*qdd[i];

    /* ---- foot contacts ---- */
    double vfree[13][6]; velocities(W, vf0, qdf, vfree);
    double fpos[4][3], fvel_free[4][3], fvel_now[4][3], nrm[4][3], gap[4], vn_min[4]; int active[4];
    for (int L=0;L<4;L++) {
        int b=body_of(L,2); double t[3];
        m3v(W->Rw[b], GO1_FOOT_OFFSET[L], t); for(int a=0;a<3;a++) fpos[L][a]=W->pw[b][a]+t[a];
        point_vel_world(W, vfree, b, GO1_FOOT_OFFSET[L], fvel_free[L]);
        point_vel_world(W, W->v, b, GO1_FOOT_OFFSET[L], fvel_now[L]);
        double h=terrain(P, fpos[L][0], fpos[L][1], nrm[L]);
        gap[L]=(fpos[L][2]-h)*nrm[L][2]-GO1_FOOT_RADIUS;
        active[L]=gap[L]<P->contact_margin;
        if (gap[L]>=0) vn_min[L]=-gap[L]/dt;
        else { vn_min[L]=P->erp*(-gap[L])/dt; if (vn_min[L]>P->max_depen_vel) vn_min[L]=P->max_depen_vel; }
        double vpre=v3dot(fvel_now[L], nrm[L]);
        if (vpre<-P->bounce_threshold && -rest*vpre>vn_min[L]) vn_min[L]=-rest*vpre;
    }
    /* Delassus matrix: response of every foot's world velocity to a unit world impulse at foot L axis k */
    double Wd[12][12], col_v0[12][6], col_qd[12][12];
    for (int L=0;L<4;L++) for (int k=0;k<3;k++) {
        double pz[13][6]; memset(pz,0,sizeof pz);
        int b=body_of(L,2); double e[3]={0,0,0}, fb[3], nb[3]; e[k]=1;
        m3tv(W->Rw[b], e, fb); v3cross(GO1_FOOT_OFFSET[L], fb, nb);
        for (int a=0;a<3;a++){ pz[b][a]=-nb[a]; pz[b][3+a]=-fb[a]; }
        double da0[6], dqd[12], dv[13][6];
        propagate(W, pz, NULL, 0, da0, dqd, dv);
        memcpy(col_v0[3*L+k], da0, sizeof da0); memcpy(col_qd[3*L+k], dqd, sizeof dqd);
        for (int M=0;M<4;M++){ double r[3]; point_vel_world(W, dv, body_of(M,2), GO1_FOOT_OFFSET[M], r);
            for (int a=0;a<3;a++) Wd[3*M+a][3*L+k]=r[a]; }
    }
    /* projected block-Jacobi (across feet) / Gauss-Seidel (within a foot) */
    double lam[4][3]; memset(lam,0,sizeof lam);
    for (int it=0; it<P->pgs_iters; it++) {
        double old[4][3]; memcpy(old, lam, sizeof lam);
        for (int L=0;L<4;L++) {
            if (!active[L]) { lam[L][0]=lam[L][1]=lam[L][2]=0; continue; }
            const double* n=nrm[L];
            double t1[3]={1-n[0]*n[0], -n[0]*n[1], -n[0]*n[2]};
            double tn=sqrt(v3dot(t1,t1)); for(int a=0;a<3;a++) t1[a]/=tn;
            double t2[3]; v3cross(n,t1,t2);
            double r[3], l[3];
            for (int a=0;a<3;a++){ r[a]=fvel_free[L][a]; l[a]=old[L][a];
                for (int M=0;M<4;M++) for(int c2=0;c2<3;c2++) r[a]+=Wd[3*L+a][3*M+c2]*old[M][c2]; }
            const double* dirs[3]={n,t1,t2};
            for (int row=0;row<3;row++) {
                const double* d=dirs[row]; double Wdv[3];
                for (int a=0;a<3;a++){ Wdv[a]=0; for(int c2=0;c2<3;c2++) Wdv[a]+=Wd[3*L+a][3*L+c2]*d[c2]; }
                double A=v3dot(d,Wdv)+P->cfm;
                double delta;
                if (row==0) { double ln=v3dot(n,l); delta=-(v3dot(d,r)-vn_min[L])/A;
                              double lnn=ln+delta; if (lnn<0) lnn=0; delta=lnn-ln; }
                else delta=-v3dot(d,r)/A;
                for (int a=0;a<3;a++){ l[a]+=delta*d[a]; r[a]+=delta*Wdv[a]; }
            }
            double ln=v3dot(n,l), lt[3]; for(int a=0;a<3;a++) lt[a]=l[a]-ln*n[a];
            double ltn=sqrt(v3dot(lt,lt));
            if (ltn>mu*ln) { double sc=(ltn>1e-12)?mu*ln/ltn:0; for(int a=0;a<3;a++) lt[a]*=sc; }
            for (int a=0;a<3;a++) lam[L][a]=ln*n[a]+lt[a];
        }
    }
    /* ---- apply impulses, integrate ---- */
    for (int L=0;L<4;L++) for(int k=0;k<3;k++) {
        double l=lam[L][k];
        for (int a=0;a<6;a++) vf0[a]+=l*col_v0[3*L+k][a];
        for (int i=0;i<12;i++) qdf[i]+=l*col_qd[3*L+k][i];
        out->contact_force[4+4*L][k]=l/dt;
    }
    for (int i=0;i<12;i++) {
        double vm=W->J[i].vmax; if (qdf[i]>vm) qdf[i]=vm; if (qdf[i]<-vm) qdf[i]=-vm;
        s->qd[i]=qdf[i]; s->q[i]+=dt*qdf[i];
    }
    m3v(R0, vf0, s->angvel); m3v(R0, vf0+3, s->linvel);
    for (int k=0;k<3;k++) s->pos[k]+=dt*s->linvel[k];
    { /* quaternion: q <- exp(dt*w_world) (x) q */
        double w[3]={s->angvel[0],s->angvel[1],s->angvel[2]};
        double wn=sqrt(v3dot(w,w)), ang=wn*dt, sc=(wn>1e-12)?sin(0.5*ang)/wn:0.5*dt, cw=cos(0.5*ang);
        double dx=w[0]*sc, dy=w[1]*sc, dz=w[2]*sc;
        double x=s->quat[0], y=s->quat[1], z=s->quat[2], ww=s->quat[3];
        double nx=cw*x+dx*ww+dy*z-dz*y, ny=cw*y-dx*z+dy*ww+dz*x, nz=cw*z+dx*y-dy*x+dz*ww, nw=cw*ww-dx*x-dy*y-dz*z;
        double inv=1.0/sqrt(nx*nx+ny*ny+nz*nz+nw*nw);
        s->quat[0]=nx*inv; s->quat[1]=ny*inv; s->quat[2]=nz*inv; s->quat[3]=nw*inv;
    }
    free(W);
}

/* Forward kinematics of the feet for the CURRENT state (used by post-physics: legged_robot.py:112-115
   reads rigid_body_state of the foot bodies after the last substep). */
void go1_oracle_feet(const Go1PhysState* s, double foot_pos[4][3], double foot_vel[4][3]) {
    Work* W=(Work*)calloc(1,sizeof(Work));
    double R0[9]; quat_to_R(s->quat, R0);
    memcpy(W->Rw[0], R0, sizeof R0); memcpy(W->pw[0], s->pos, sizeof s->pos);
    double v0[6]; m3tv(R0, s->angvel, v0); m3tv(R0, s->linvel, v0+3);
    for (int L=0;L<4;L++) for (int j=0;j<3;j++) {
        int i=3*L+j, b=body_of(L,j), p=parent_of(L,j);
        leg_joint(L,j,&W->J[i]);
        double Rj[9]; axis_R(W->J[i].axis, s->q[i], Rj);
        for (int a=0;a<3;a++) for(int c2=0;c2<3;c2++) W->E[i][3*a+c2]=Rj[3*c2+a];
        m3m(W->Rw[p], Rj, W->Rw[b]);
        double t[3]; m3v(W->Rw[p], W->J[i].r, t);
        for (int k=0;k<3;k++) W->pw[b][k]=W->pw[p][k]+t[k];
    }
    double v[13][6]; velocities(W, v0, s->qd, v);
    for (int L=0;L<4;L++) { int b=body_of(L,2); double t[3];
        m3v(W->Rw[b], GO1_FOOT_OFFSET[L], t); for(int a=0;a<3;a++) foot_pos[L][a]=W->pw[b][a]+t[a];
        point_vel_world(W, v, b, GO1_FOOT_OFFSET[L], foot_vel[L]); }
    free(W);
}

/* Generalised-acceleration probe for the invariant tests: contact-free ABA result in body coordinates
   (a0 = spatial acceleration of the base with gravity folded in as a' = a - a_g) */
void go1_oracle_aba(const Go1PhysParams* P, const Go1PhysDR* dr, const Go1PhysState* s,
                    const double tau[12], double a0_out[6], double qdd_out[12]) {
    Go1PhysParams Q=*P; Q.contact_margin=-1e9; Q.pgs_iters=0;   /* disable foot contacts */
    for (int k=0;k<4;k++){ Q.pen_k[k]=0; Q.pen_c[k]=0; }
    Q.limit_k=0; Q.limit_c=0;
    Go1PhysState t=*s; Go1PhysOut o;
    /* finite-difference-free: rerun the substep with the probe settings and recover accelerations from
       the velocity change (exact for the semi-implicit Euler update used here). */
    double R0[9]; quat_to_R(s->quat, R0);
    double v0[6]; m3tv(R0, s->angvel, v0); m3tv(R0, s->linvel, v0+3);
    go1_oracle_substep(&Q, dr, &t, tau, &o);
    double v1[6]; m3tv(R0, t.angvel, v1); m3tv(R0, t.linvel, v1+3);
    for (int k=0;k<6;k++) a0_out[k]=(v1[k]-v0[k])/Q.dt;
    for (int i=0;i<12;i++) qdd_out[i]=(t.qd[i]-s->qd[i])/Q.dt;
}

typedef struct { const Go1PhysParams* P; const Go1PhysDR* dr; Go1PhysState* s; const double* tau; Go1PhysOut* out; double* fp; double* fv; int e0, e1; } BatchJob;
static int g_max_threads = 0;        /* 0 = all online processors */
void go1_oracle_set_threads(int t) { g_max_threads = t > 0 ? t : 0; }
static int n_threads(int n) {
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    int T = (int)(c > 0 ? c : 1);
    if (g_max_threads > 0 && T > g_max_threads) T = g_max_threads;
    if (T > 256) T = 256;
    if (T > n / 4) T = n / 4;
    return T < 1 ? 1 : T;
}
static void* substep_job(void* a) {
    BatchJob* j = (BatchJob*)a;
    for (int e=j->e0;e<j->e1;e++) go1_oracle_substep(j->P, j->dr+e, j->s+e, j->tau+12*e, j->out+e);
    return NULL;
}
static void* feet_job(void* a) {
    BatchJob* j = (BatchJob*)a;
    for (int e=j->e0;e<j->e1;e++) go1_oracle_feet(j->s+e, (double(*)[3])(j->fp+12*e), (double(*)[3])(j->fv+12*e));
    return NULL;
}
void go1_oracle_feet_batch(int n, const Go1PhysState* s, double* foot_pos, double* foot_vel) {
    BatchJob jobs[256]; pthread_t th[256];
    int T = n_threads(n);
    for (int t=0;t<T;t++) {
        jobs[t] = (BatchJob){NULL, NULL, (Go1PhysState*)s, NULL, NULL, foot_pos, foot_vel, t*n/T, (t+1)*n/T};
        if (t < T-1) pthread_create(&th[t], NULL, feet_job, &jobs[t]);
    }
    feet_job(&jobs[T-1]);
    for (int t=0;t<T-1;t++) pthread_join(th[t], NULL);
}

void go1_oracle_substep_batch(const Go1PhysParams* P, int n, const Go1PhysDR* dr, Go1PhysState* s,
                              const double* tau, Go1PhysOut* out) {
    /* envs are independent: split the batch over all host cores (pthreads; this image has no OpenMP runtime) */
    BatchJob jobs[256]; pthread_t th[256];
    int T = n_threads(n);
    for (int t=0;t<T;t++) {
        jobs[t] = (BatchJob){P, dr, s, tau, out, NULL, NULL, t*n/T, (t+1)*n/T};
        if (t < T-1) pthread_create(&th[t], NULL, substep_job, &jobs[t]);
    }
    substep_job(&jobs[T-1]);
    for (int t=0;t<T-1;t++) pthread_join(th[t], NULL);
}
