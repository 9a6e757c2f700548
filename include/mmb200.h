/*
 * mmb200.h — C ABI of libmmb200.so: batched material-point updates on NVIDIA B200 (sm_100a).
 *
 * One call = one pass of `material_response` over N independent material (quadrature) points.
 * The reference (kimauth/MaterialModels.jl) has NO batched entry point and NO FFI: its boundary is
 * the Julia generic
 *     material_response(m, strain, state, Δt; cache, options) -> (stress, tangent, new_state)
 * (reference src/MaterialModels.jl:41-52), called once per point by the FE assembler.  Each
 * function below replaces the loop over points around one method of that generic; the reference
 * method it replaces is cited next to it.  INTEGRATION.md shows the Julia `ccall` stubs.
 *
 * Conventions (identical to Tensors.jl storage, so Julia arrays reinterpret zero-copy):
 *   SymmetricTensor{2,3}: 6 doubles  (11,21,31,22,32,33)
 *   Tensor{2,3}         : 9 doubles  column-major (11,21,31,12,22,32,13,23,33)
 *   SymmetricTensor{4,3}: 36 doubles, index I + 6*J, I<->(ij), J<->(kl) in the pair order above
 *   Tensor{1,dim}       : dim doubles, LAST component = normal separation (XuNeedleman)
 * Layout of every per-point array with `nc` components:
 *   MM_LAYOUT_SOA: a[c*N + i]   (component-major; the fast path — fully coalesced)
 *   MM_LAYOUT_AOS: a[i*nc + c]  (== reinterpret(Float64, ::Vector{SymmetricTensor{2,3,Float64,6}}))
 * State packing: Plastic  [εᵖ(6), κ, α(6), μ]            = 14 doubles (src/Plastic.jl:41-46 field order)
 *                Crystal  [σ(6), κ(S), α(S), μ(S)], S=12 = 42 doubles (CrystalViscoPlastic.jl:52-57)
 *
 * All array pointers of the mm_* entry points are DEVICE pointers owned by the caller; nothing is
 * allocated on the hot path; `stream` is a cudaStream_t (NULL = default stream); calls are
 * asynchronous with respect to the host.  The mmh_* entry points take HOST pointers and run the
 * same kernels through a pipelined H2D -> kernel -> D2H chunk loop (blocking).
 * Return value: 0 ok, <0 error (see mm_last_error()).  Non-convergence of a local Newton solve is
 * never an abort: it is counted in *n_fail and flagged per point; the host mirror raises the
 * reference's error text ("Material model not converged. Could not find material state.",
 * src/Plastic.jl:157, CrystalViscoPlastic.jl:92) when the count is non-zero.
 * Non-finite inputs follow the reference's branches (`if Φ <= 0 … else` sends a NaN Φ to the Newton branch, ⟨Φ⟩ = (Φ+|Φ|)/2
 * propagates it): such points end flagged as failed, their neighbours are unaffected.  Newton iteration counts (`iters`) are
 * uint16 (NLsolve's default cap is 1000; they saturate at 65535); the wrappers' outer counts are uint8 (default cap 10).
 */
#ifndef MMB200_H
#define MMB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM_LAYOUT_SOA 0
#define MM_LAYOUT_AOS 1

#define MM_NSLIP 12 /* FCC 4x3 and BCC 6x2 both give 12 systems (slipsystems.jl:9-35) */

#define MM_OK 0
#define MM_ERR_ARG (-1)
#define MM_ERR_CUDA (-2)
#define MM_ERR_NOGPU (-3)
#define MM_ERR_CAPACITY (-4) /* compact return: more points yielded than the caller's row buffer holds */

/* hyperelastic models (mm_hyper `model`) */
#define MM_NEOHOOK 0  /* src/FiniteStrain/neohook.jl:35-43  */
#define MM_YEOH 1     /* src/FiniteStrain/yeoh.jl:37-54     */
#define MM_STVENANT 2 /* src/FiniteStrain/stvenant.jl:36-48 */
/* strain measure handed to mm_hyper (`strain_kind`) */
#define MM_STRAIN_C 0 /* RightCauchyGreen C, 6 comps (native, neohook.jl:35)                       */
#define MM_STRAIN_F 1 /* DeformationGradient F, 9 comps; C = tdot(F) = FᵀF (src/traits.jl:35-38) */
#define MM_STRAIN_E 2 /* GreenLagrange E, 6 comps; C = 2E + I (src/traits.jl:43-46)              */
/* stress / tangent measure returned by mm_hyper_ex (`tangent_kind`), src/traits.jl:14-26, :60-77 */
#define MM_TANGENT_DSDC 0 /* S (6),  ∂S∂C (36)  — native                                          */
#define MM_TANGENT_DSDE 1 /* S (6),  ∂S∂E = 2 ∂S∂C (36)                       traits.jl:61-63    */
#define MM_TANGENT_DPDF 2 /* Pᵀ = S⋅Fᵀ (9, Tensor{2,3}), ∂Pᵀ∂F (81, Tensor{4,3}: i+3j+9k+27l); needs MM_STRAIN_F   traits.jl:67-77 */
/* crystal structures (slipsystems.jl:3-4) */
#define MM_FCC 0
#define MM_BCC 1

/* per-point flag bits written by mm_plastic */
#define MM_FLAG_PLASTIC 0x01 /* Φ_trial > 0: the Newton branch was taken (Plastic.jl:134)   */
#define MM_FLAG_FAILED 0x80  /* local Newton did not converge (Plastic.jl:157)              */
/* mm_crystal `active` is a bit mask: bit i set <=> Φ_i > 0 at the solution; bit 15 = failed */
#define MM_ACTIVE_FAILED 0x8000

typedef struct MMLinearElastic { double E, nu; } MMLinearElastic;                 /* LinearElastic.jl:10-19 */
typedef struct MMPlastic { double E, nu, sigma_y, H, r, kappa_inf, alpha_inf; } MMPlastic; /* Plastic.jl:15-36 */
typedef struct MMHyper { double lambda, mu, c2, c3; } MMHyper;                   /* neohook.jl:10-13, yeoh.jl:11-16 */
typedef struct MMXuNeedleman { double Phi_n, Phi_t, delta_n, delta_t, Delta_n_star; } MMXuNeedleman; /* XuNeedleman.jl:22-28 */
typedef struct MMCrystal {                                                        /* CrystalViscoPlastic.jl:18-45 */
    double E, nu, tau_y, H_iso, H_kin, q, alpha_inf, t_star, sigma_c, m;
    int32_t nslip;  /* must be MM_NSLIP */
    int32_t pad_;
    double MS[MM_NSLIP][9];        /* m_i ⊗ s_i, Tensor{2,3} column-major (:42)                    */
    double H[MM_NSLIP][MM_NSLIP];  /* H[i][j] = Julia H[i+1,j+1] = H_iso (q + (1-q) δ_ij)   (:41)  */
} MMCrystal;
typedef struct MMTransverselyIsotropic { double L_perp, L_par, M_par, G_perp, G_par; } MMTransverselyIsotropic; /* transverselyisotropic.jl:19-37 */
/* Newton options.  Plastic: NLsolve `ftol` on max|F| and `iterations` (defaults 1e-8 / 1000).
 * Crystal: `solve(...; tol, max_iter)` on ||r||_2, strict < (nonlinear_solver.jl:52; 1e-8 / 100). */
typedef struct MMNewtonOpts { double tol; int32_t max_iter; int32_t pad_; } MMNewtonOpts;

/* ---- library ------------------------------------------------------------------------------- */
const char* mm_version(void);
const char* mm_last_error(void);
int mm_device_count(void);

/* test / tuning hook: A-B switches of the crystal kernels ("crystal_force_general", "crystal_thresh", "crystal_chunk",
 * "crystal_fixup_per_sm"; "crystal_finish_prefetch" = points ahead for which the finish kernel asks L2 for its SoA input lines, 0 = no hint, default 4096; "crystal_variant" 0 = dense pass, 1 = sparse per-lane lists, 2 = phased, 3 / 4 = sparse / dense over chunks re-ordered by
 * first-pass candidate count, 5 / 6 = active-set pass with / without that re-ordering, -1 = default: 6 (SoA) or 5 (AoS)) and of the dimension wrappers ("wrapped_j2_variant" 0 = one thread per point, 1 = state
 * machine, 2 = pass-granular trips; "wrapped_j2_th_outer", "wrapped_j2_th_final", "wrapped_j2_chunk"; "wrapped_crystal_variant" 0 = one
 * thread per point, 1 = rounds; "wrapped_crystal_slab") and of the XuNeedleman SoA path ("xu_soa_tiled" 0 = direct kernel, 1 = copy-engine
 * tiles), of the J2 SoA kernel ("plastic_soa_prefetch" = points ahead for which the input lines are requested into L2, 0 = no hint, default 28416; "stream_soa_prefetch" the same for the
 * LinearElastic / hyperelastic / XuNeedleman SoA kernels, -1 = their own defaults: hints, no value changes) and of the host-buffer J2 call ("mmh_plastic_via_factored" 1 = dense tangent rebuilt on the host from factored rows for batches of >= 2^18 points,
 * 2 = for any size, 0 = plain dense transfer); initial values come from the MM_CRYSTAL_* / MM_WRAPPED_* environment variables, read once per process.
 * Returns the previous value, -1 for an unknown name.  Results do not depend on the scheduling switches (chunk, thresholds, re-ordering, slabs: bit for bit);
 * the Newton pass variants of the crystal update solve the same Newton step by different algebra and agree to rounding — active masks and iteration counts
 * identical, values within 1e-12 (tests/test_gpu_parity.py runs every variant against the oracle). */
int mm_set_tuning(const char* name, int value);

/* ---- host-side parameter helpers (run once per material, like the Julia constructors) ------ */
/* elastic_tangent_3D(E, ν) — src/LinearElastic.jl:27-35 */
int mm_elastic_tangent_3d(double E, double nu, double* out36);
/* slipsystems(FCC()/BCC(), RodriguesParam(x,y,z)) — slipsystems.jl:37-53; planes/dirs: [12][3] */
int mm_slipsystems(int structure, const double rodrigues[3], double* planes, double* dirs);
/* CrystalViscoPlastic(...) constructor — fills nslip, MS and H from the scalars already in *p (:38-45) */
int mm_crystal_setup(MMCrystal* p, const double* planes, const double* dirs);
/* XuNeedleman(; σₘₐₓ, τₘₐₓ, Φₙ, Φₜ, Δₙˢ) — XuNeedleman.jl:30-34 */
int mm_xu_needleman_setup(double sigma_max, double tau_max, double Phi_n, double Phi_t, double Delta_n_star, MMXuNeedleman* out);

/* ---- the hot path: device-pointer entry points --------------------------------------------- */
/* material_response(::LinearElastic, ε) — src/LinearElastic.jl:57-60.
 * eps[6], sig[6]; tan[36] may be NULL (the reference returns the one shared m.Eᵉ; use mm_elastic_tangent_3d). */
int mm_linear_elastic(const MMLinearElastic* p, const double* eps, double* sig, double* tan_or_null,
                      int64_t N, int layout, void* stream);

/* material_response(::Plastic, ε, state; cache, options) — src/Plastic.jl:126-160 (+ residuals :162-172,
 * NLsolve newton at :146-148).  eps[6] TOTAL strain, state_in/out[14] (may alias), sig[6], tan[36] (not
 * major-symmetric in general; may be NULL), flags (uint8) / iters (uint16) (may be NULL), *n_fail += #non-converged. */
int mm_plastic(const MMPlastic* p, const double* eps, const double* state_in, double* sig, double* tan,
               double* state_out, uint8_t* flags, uint16_t* iters, int32_t* n_fail, const MMNewtonOpts* opts,
               int64_t N, int layout, void* stream);

/* ---- compact return of the J2 update (opt-in) — reference src/Plastic.jl:134-135 --------------------------------------------
 * The reference hands every ELASTIC point the one stored m.Eᵉ and its unchanged state object; only a point that took the Newton
 * branch gets a tangent and a state of its own.  mm_plastic writes 36 + 14 doubles for every point all the same (400 of its 608 bytes
 * per point).  The compact form returns σ[6], flags and iters for all N points and ONE ROW per yielded point (MM_FLAG_PLASTIC), rows in
 * point order (row index of point i = number of yielded points before i):
 *     MM_ROWS_TANGENT_STATE   [∂σ∂ε(36) | new state(14)]        50 doubles      MM_ROWS_TANGENT   [∂σ∂ε(36)]
 *     MM_ROWS_FACTORED_STATE  [gξ, u(6), n(6) | new state(14)]  27 doubles      MM_ROWS_FACTORED  [gξ, u(6), n(6)]
 * with ∂σ∂ε[I + 6J] = fma(u[I], n[J], fma(−gξ, I_dev[I,J], Eᵉ[I,J])) — 13 numbers carry the consistent tangent exactly.
 * Elastic points: tangent = mm_elastic_tangent_3d(E, ν), state = the caller's state_in.  *n_rows (DEVICE int64 for mm_, host for mmh_)
 * receives the number of yielded points; rows beyond rows_capacity are not written (mmh_: MM_ERR_CAPACITY, everything else complete).
 * state_out (dense, may be NULL, may alias state_in) is written for all points when given.  `workspace`: device scratch of
 * mm_plastic_compact_workspace(N, state_out != NULL) bytes (nothing is allocated on the hot path). */
#define MM_ROWS_TANGENT_STATE 0
#define MM_ROWS_FACTORED_STATE 1
#define MM_ROWS_TANGENT 2
#define MM_ROWS_FACTORED 3
int mm_plastic_row_width(int row_kind);
uint64_t mm_plastic_compact_workspace(int64_t N, int want_state_out);
int mm_plastic_compact(const MMPlastic* p, const double* eps, const double* state_in, double* sig, uint8_t* flags, uint16_t* iters, int row_kind,
                       double* rows, int64_t rows_capacity, int64_t* n_rows, double* state_out, int32_t* n_fail, const MMNewtonOpts* opts,
                       void* workspace, uint64_t workspace_bytes, int64_t N, int layout, void* stream);

/* material_response(::NeoHook/::Yeoh/::StVenant, C) — neohook.jl:35-43, yeoh.jl:37-54, stvenant.jl:36-48, and the
 * trait route material_response(∂S∂C(), m, DeformationGradient(F), ...) — src/traits.jl:91-130.
 * strain[6 or 9], S[6] (2nd Piola–Kirchhoff), dSdC[36] (may be NULL). */
int mm_hyper(int model, const MMHyper* p, const double* strain, int strain_kind, double* S, double* dSdC,
             int64_t N, int layout, void* stream);

/* material_response(output_tangent, m, strain::StrainMeasure, state, Δt) — the full trait layer src/traits.jl:91-107
 * (SURVEY §8(f) rank 2): any strain kind in, any tangent kind out, fused into the hyperelastic kernel. */
int mm_hyper_ex(int model, const MMHyper* p, const double* strain, int strain_kind, int tangent_kind, double* stress, double* tangent,
                int64_t N, int layout, void* stream);

/* material_response(::CrystalViscoPlastic{12}, Δε, state, Δt; options) — CrystalViscoPlastic.jl:64-94
 * (+ residuals :97-115, get_state_vars :118-132, solve nonlinear_solver.jl:52-62).
 * deps[6] strain INCREMENT, state_in/out[42] (may alias), sig[6], tan[36], active (uint16 mask), iters (uint16). */
int mm_crystal(const MMCrystal* p, const double* deps, double dt, const double* state_in, double* sig, double* tan,
               double* state_out, uint16_t* active, uint16_t* iters, int32_t* n_fail, const MMNewtonOpts* opts,
               int64_t N, int layout, void* stream);

/* material_response(::XuNeedleman, Δ::Tensor{1,dim}) — XuNeedleman.jl:75-91.  dim = 2 or 3.
 * jump[dim], T[dim], dTdD[dim*dim] column-major (symmetric Hessian stored in full, as Tensor{2,dim}). */
int mm_xu_needleman(const MMXuNeedleman* p, const double* jump, double* T, double* dTdD, int dim,
                    int64_t N, int layout, void* stream);

/* ---- dimension wrappers — reference src/wrappers.jl ----------------------------------
 * material_response(dim::AbstractDim, m, Δε_dim, state, Δt; cache, options): strain-restricted kinds pad the strain with
 * zeros and cut the 3D result down (wrappers.jl:29-41); stress-restricted kinds iterate the out-of-plane strains until
 * the out-of-plane stresses vanish (options :plane_stress_tol = 1e-8 on the Mandel 2-norm, :plane_stress_max_iter = 10)
 * and return the Schur-complement tangent (wrappers.jl:45-79).  Reduced storage = Tensors.jl SymmetricTensor{2,d}/{4,d}:
 * nc = mm_wrap_ncomp(kind) = 1, 3 (11,21,22), or 6; tangent nc*nc with index I + nc*J.  MM_DIM_SHELL_STRESS keeps the
 * full 3D storage with zero 33-rows/columns (wrappers.jl:131-146).
 * flags (Plastic): MM_FLAG_PLASTIC / MM_FLAG_FAILED of the LAST inner evaluation, MM_FLAG_WRAPPER_FAILED when the outer
 * Newton did not converge ("Not converged. Could not find plane stress state.", wrappers.jl:78). */
#define MM_DIM_UNIAXIAL_STRAIN 1 /* UniaxialStrain{1}  wrappers.jl:4  */
#define MM_DIM_PLANE_STRAIN 2    /* PlaneStrain{2}     wrappers.jl:6  */
#define MM_DIM_UNIAXIAL_STRESS 3 /* UniaxialStress{1}  wrappers.jl:5  */
#define MM_DIM_PLANE_STRESS 4    /* PlaneStress{2}     wrappers.jl:7  */
#define MM_DIM_SHELL_STRESS 5    /* ShellStress{3}     wrappers.jl:8  */
#define MM_FLAG_WRAPPER_FAILED 0x40
int mm_wrap_ncomp(int dim_kind);
int mm_wrapped_linear_elastic(int dim_kind, const MMLinearElastic* p, const double* eps_red, double* sig_red, double* tan_red,
                              uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* wrapper_opts, int64_t N, int layout, void* stream);
int mm_wrapped_plastic(int dim_kind, const MMPlastic* p, const double* eps_red, const double* state_in, double* sig_red, double* tan_red,
                       double* state_out, uint8_t* flags, uint16_t* iters, uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* opts,
                       const MMNewtonOpts* wrapper_opts, int64_t N, int layout, void* stream);
/* ... and around CrystalViscoPlastic{12}: deps_red is the reduced strain INCREMENT, state 42 per point; `active` as mm_crystal plus
 * bit 13 (0x2000) = the outer Newton did not converge.  The stress-restricted kinds run as rounds of mm_crystal's kernels on the
 * still-unconverged points (stream-ordered, no host synchronisation; workspace of ~1.1 KB per point, at most 2^21 points at a time,
 * from the library's own stream-ordered memory pool).  state_out may alias state_in. */
int mm_wrapped_crystal(int dim_kind, const MMCrystal* p, const double* deps_red, double dt, const double* state_in, double* sig_red,
                       double* tan_red, double* state_out, uint16_t* active, uint16_t* iters, uint8_t* outer_iters, int32_t* n_fail,
                       const MMNewtonOpts* opts, const MMNewtonOpts* wrapper_opts, int64_t N, int layout, void* stream);
/* the same wrappers around TransverselyIsotropic (wrappers.jl is generic in the material); a3[3] per point is the state */
int mm_wrapped_transversely_isotropic(int dim_kind, const MMTransverselyIsotropic* p, const double* eps_red, const double* a3,
                                      double* sig_red, double* tan_red, uint8_t* outer_iters, int32_t* n_fail,
                                      const MMNewtonOpts* wrapper_opts, int64_t N, int layout, void* stream);

/* ---- TransverselyIsotropic — reference src/transverselyisotropic.jl -----------------------
 * TransverselyIsotropic(; E_L, E_T, G_LT, ν_LT, ν_TT) (:39-53) and material_response(m, ε, state) (:92-106): the stiffness is
 * rebuilt per point from the unit direction a3 held in the state.  eps[6], a3[3] (the state; returned unchanged), sig[6],
 * tan[36] (may be NULL). */
int mm_transversely_isotropic_setup(double E_L, double E_T, double G_LT, double nu_LT, double nu_TT, MMTransverselyIsotropic* out);
int mm_transversely_isotropic(const MMTransverselyIsotropic* p, const double* eps, const double* a3, double* sig, double* tan,
                              int64_t N, int layout, void* stream);

/* ---- elastic_strain_energy_density — LinearElastic.jl:63, neohook.jl:29-33, yeoh.jl:31-35,
 * stvenant.jl:28-34, and through any strain measure (traits.jl:133-136).  psi[N] (one double per point, layout-free). */
int mm_linear_elastic_energy(const MMLinearElastic* p, const double* eps, double* psi, int64_t N, int layout, void* stream);
int mm_hyper_energy(int model, const MMHyper* p, const double* strain, int strain_kind, double* psi, int64_t N, int layout, void* stream);

/* ---- the caller side fused around the update — the quadrature loop of a small-strain solid element,
 * i.e. what a Ferrite.jl-style assemble_cell! does around the reference's per-point material_response (src/MaterialModels.jl:41-52):
 *     ε = sym(Σ_a u_a ⊗ ∇N_a);  σ, D, state_new = material_response(m, ε, state);  fe[a] += (σ ⋅ ∇N_a) dΩ
 * nn nodes and nq quadrature points per cell, supported (nn, nq): (4,1) (4,4) (8,1) (8,8) (10,4).  P = ncell*nq points, point
 * index = cell*nq + q.  All arrays component-major: per point dNdx[(3a+k)*P + i] (∇N_a in physical coordinates, as CellValues
 * holds it after reinit!), dOmega[i], state / sig / tan [c*P + i]; per cell ue, fe [(3a+k)*ncell + cell].
 * sig (6 per point) and tan (36 per point) may be NULL: explicit / matrix-free codes need neither, and leaving the tangent out
 * removes 288 of the 608 algorithmic bytes per point of the unfused J2 update.
 * ke (may be NULL): the element stiffness ke[3a+i][3b+j] = Σ_q δε_(a,i) ⊡ D ⊡ Δε_(b,j) dΩ, (3nn)² per cell, stored
 * ke[((3a+i)*3nn + 3b+j)*ncell + cell]; D stays in registers (not major-symmetric for J2 with kinematic hardening, so all
 * (3nn)² entries are written). */
int mm_cells_linear_elastic(const MMLinearElastic* p, int nn, int nq, const double* dNdx, const double* dOmega, const double* ue,
                            double* fe, double* ke, double* sig, double* tan, int64_t ncell, void* stream);
int mm_cells_plastic(const MMPlastic* p, int nn, int nq, const double* dNdx, const double* dOmega, const double* ue,
                     const double* state_in, double* fe, double* ke, double* sig, double* tan, double* state_out, uint8_t* flags,
                     uint16_t* iters, int32_t* n_fail, const MMNewtonOpts* opts, int64_t ncell, void* stream);

int mm_cells_transversely_isotropic(const MMTransverselyIsotropic* p, int nn, int nq, const double* dNdx, const double* dOmega,
                                    const double* ue, const double* a3 /* [3][P] unit directions */, double* fe, double* ke, double* sig,
                                    double* tan, int64_t ncell, void* stream);
/* total-Lagrangian counterpart for NeoHook / Yeoh / StVenant: F = I + Σ_a u_a ⊗ ∇N_a (∇N w.r.t. the reference configuration),
 * (Pᵀ, ∂Pᵀ∂F) through the trait route (src/traits.jl:67-77), fe[a][i] = Σ_q Σ_j P_ij ∇N_a[j] dΩ.  Pt (9 per point, i + 3j) and
 * dPdF (81 per point) may be NULL. */
int mm_cells_hyper(int model, const MMHyper* p, int nn, int nq, const double* dNdx, const double* dOmega, const double* ue,
                   double* fe, double* Pt, double* dPdF, int64_t ncell, void* stream);

/* Deterministic synthetic inputs (bench + parity harness): out(i,c) = lo[c] + (hi[c]-lo[c]) * u(i,c),
 * u(i,c) = (splitmix64(seed + i*nc + c) >> 11) * 2^-53.  lo/hi are HOST arrays of nc doubles. */
int mm_synth_uniform(double* out, int64_t N, int nc, int layout, uint64_t seed, const double* lo, const double* hi,
                     void* stream);

/* ---- host-buffer entry points (same arguments, HOST pointers; pinned memory is fastest) ----- */
void* mmh_alloc_pinned(uint64_t bytes);
void mmh_free_pinned(void* p);
int mmh_linear_elastic(const MMLinearElastic* p, const double* eps, double* sig, double* tan_or_null, int64_t N, int layout);
int mmh_plastic(const MMPlastic* p, const double* eps, const double* state_in, double* sig, double* tan, double* state_out,
                uint8_t* flags, uint16_t* iters, int32_t* n_fail, const MMNewtonOpts* opts, int64_t N, int layout);
/* compact return through the chunk pipeline: per point 160 B up and 51 B + (yielded fraction) x 8 x row width down instead of 450 B.
 * flags is required.  mmh_plastic_compact_resident keeps the state in DEVICE memory (d_state: one N-point array in `layout`, updated in
 * place; row kind MM_ROWS_TANGENT or MM_ROWS_FACTORED): only ε goes up and only σ, flags, iters and the tangent rows come down. */
int mmh_plastic_compact(const MMPlastic* p, const double* eps, const double* state_in, double* sig, uint8_t* flags, uint16_t* iters, int row_kind,
                        double* rows, int64_t rows_capacity, int64_t* n_rows, double* state_out, int32_t* n_fail, const MMNewtonOpts* opts,
                        int64_t N, int layout);
int mmh_plastic_compact_resident(const MMPlastic* p, const double* eps, double* d_state, double* sig, uint8_t* flags, uint16_t* iters, int row_kind,
                                 double* rows, int64_t rows_capacity, int64_t* n_rows, int32_t* n_fail, const MMNewtonOpts* opts, int64_t N, int layout);
/* dense arrays from the compact return (host threads): tan[36] and / or state_out[14] for all N points in `layout`, bit-identical to what
 * mmh_plastic writes.  state_out needs a ..._STATE row kind and state_in; state_out == state_in updates the caller's state in place
 * (only the yielded points are touched).  mmh_set_host_threads(n): threads used by the expander (0 = all cores, capped at 64). */
int mmh_plastic_expand(const MMPlastic* p, int row_kind, const uint8_t* flags, const double* rows, int64_t n_rows, const double* state_in,
                       double* tan, double* state_out, int64_t N, int layout);
int mmh_set_host_threads(int n);
int mmh_hyper(int model, const MMHyper* p, const double* strain, int strain_kind, double* S, double* dSdC, int64_t N, int layout);
int mmh_crystal(const MMCrystal* p, const double* deps, double dt, const double* state_in, double* sig, double* tan,
                double* state_out, uint16_t* active, uint16_t* iters, int32_t* n_fail, const MMNewtonOpts* opts, int64_t N, int layout);
/* ... with the state RESIDENT in device memory (one N-point array in `layout`, updated in place: what a time-stepping caller keeps between
 * steps): only Δε goes up (48 B per point) and σ, ∂σ∂ε, active, iters come down — the 2 x 336 B of state never cross the link */
int mmh_crystal_resident(const MMCrystal* p, const double* deps, double dt, double* d_state, double* sig, double* tan, uint16_t* active,
                         uint16_t* iters, int32_t* n_fail, const MMNewtonOpts* opts, int64_t N, int layout);
int mmh_xu_needleman(const MMXuNeedleman* p, const double* jump, double* T, double* dTdD, int dim, int64_t N, int layout);
int mmh_wrapped_linear_elastic(int dim_kind, const MMLinearElastic* p, const double* eps_red, double* sig_red, double* tan_red,
                               uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* wrapper_opts, int64_t N, int layout);
int mmh_wrapped_plastic(int dim_kind, const MMPlastic* p, const double* eps_red, const double* state_in, double* sig_red, double* tan_red,
                        double* state_out, uint8_t* flags, uint16_t* iters, uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* opts,
                        const MMNewtonOpts* wrapper_opts, int64_t N, int layout);
int mmh_wrapped_crystal(int dim_kind, const MMCrystal* p, const double* deps_red, double dt, const double* state_in, double* sig_red,
                        double* tan_red, double* state_out, uint16_t* active, uint16_t* iters, uint8_t* outer_iters, int32_t* n_fail,
                        const MMNewtonOpts* opts, const MMNewtonOpts* wrapper_opts, int64_t N, int layout);
int mmh_wrapped_transversely_isotropic(int dim_kind, const MMTransverselyIsotropic* p, const double* eps_red, const double* a3,
                                       double* sig_red, double* tan_red, uint8_t* outer_iters, int32_t* n_fail,
                                       const MMNewtonOpts* wrapper_opts, int64_t N, int layout);
int mmh_hyper_ex(int model, const MMHyper* p, const double* strain, int strain_kind, int tangent_kind, double* stress, double* tangent,
                 int64_t N, int layout);
int mmh_transversely_isotropic(const MMTransverselyIsotropic* p, const double* eps, const double* a3, double* sig, double* tan,
                               int64_t N, int layout);
int mmh_linear_elastic_energy(const MMLinearElastic* p, const double* eps, double* psi, int64_t N, int layout);
int mmh_hyper_energy(int model, const MMHyper* p, const double* strain, int strain_kind, double* psi, int64_t N, int layout);
/* points per pipeline chunk of the mmh_* calls (default 1<<20); returns the previous value */
int64_t mmh_set_chunk_points(int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* MMB200_H */
