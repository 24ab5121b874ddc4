/*
 * b200fem.h -- C ABI of the B200-native element-integration engine.
 *
 * Scope: the hot path behind nutils' Sample.integrate / Sample.integral /
 * Topology.integral -> sparse assembly (reference: evalf/nutils @ 37d1cc5).
 * The reference has NO native interface on this path (it is pure Python and
 * its only extension points are Python-level), so every
 * entry point below cites the reference code whose work it replaces; the
 * conventions follow the reference's one ctypes boundary, the MKL backend
 * (src/nutils/matrix/_mkl.py:10-97): plain C, caller-owned C-contiguous
 * buffers, integer status return, no callbacks, no exceptions.
 *
 * All functions return B2_OK (0) or a negative B2_E* code; b2_strerror()
 * gives a message, b2_last_error(ctx) the detail of the last failure.
 * A context is bound to one CUDA device and is not fork-safe (the reference
 * forks in nutils.parallel: create contexts after forking).
 * Pointers named *_host are host memory, *_dev are device memory on the
 * context's device.  Index arrays on the ABI are int64 like the reference's
 * (function.as_csr returns int64 rowptr/colidx), values are float64.
 */
#ifndef B200FEM_H
#define B200FEM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define B2_OK 0
#define B2_EINVAL -1      /* invalid argument */
#define B2_ENOMEM -2      /* host or device allocation failed */
#define B2_ECUDA -3       /* CUDA runtime error (see b2_last_error) */
#define B2_ENODEV -4      /* no usable sm_100 device */
#define B2_EUNSUPPORTED -5 /* valid request outside the implemented path */
#define B2_ENOTCONVERGED -6 /* iterative solve stopped above the requested tolerance (cf. nutils.matrix.ToleranceNotReached) */

#define B2_MAX_DIMS 3
#define B2_MAX_DEGREE 4
#define B2_MAX_FORMS 4

typedef struct b2_ctx b2_ctx;
typedef struct b2_basis b2_basis;
typedef struct b2_quad b2_quad;
typedef struct b2_geom b2_geom;
typedef struct b2_pattern b2_pattern;

const char* b2_strerror(int code);
const char* b2_last_error(const b2_ctx* ctx);
int b2_version(void);

/* ---- context ---------------------------------------------------------------------------------
 * Replaces the process-level state of the reference's loop runner: nutils.parallel.fork /
 * ctxrange (src/nutils/parallel.py:27-154) -- one context per process and GPU. */
int b2_device_count(int* count);
int b2_ctx_create(int device, b2_ctx** out);
int b2_ctx_destroy(b2_ctx* ctx);
/* run on an externally owned CUDA stream (cudaStream_t as void*; NULL = the context's own stream) */
int b2_ctx_set_stream(b2_ctx* ctx, void* cuda_stream);
int b2_ctx_synchronize(b2_ctx* ctx);
/* CUDA-event timing on the context's stream (used by bench.py for the roofline figure) */
int b2_ctx_timer_start(b2_ctx* ctx);
int b2_ctx_timer_stop(b2_ctx* ctx, float* milliseconds);
/* number of kernels launched by this context since creation */
int64_t b2_ctx_launch_count(const b2_ctx* ctx);
/* per-kernel device time: after b2_ctx_set_option(ctx, "time_kernels", 1) every assembly kernel launch is
 * bracketed by CUDA events on the context's stream; this call synchronises, returns the summed duration and
 * the number of launches since the last call, and resets both (bench.py: roofline.achieved) */
int b2_ctx_kernel_time(b2_ctx* ctx, double* milliseconds, int64_t* launches);
/* pinned host memory for the end-to-end path */
int b2_host_alloc(b2_ctx* ctx, int64_t nbytes, void** out_host);
int b2_host_free(b2_ctx* ctx, void* host);
int b2_device_alloc(b2_ctx* ctx, int64_t nbytes, void** out_dev);
int b2_device_free(b2_ctx* ctx, void* dev);
int b2_memcpy_d2h(b2_ctx* ctx, void* dst_host, const void* src_dev, int64_t nbytes);
int b2_memcpy_h2d(b2_ctx* ctx, void* dst_dev, const void* src_host, int64_t nbytes);
int b2_memset_zero(b2_ctx* ctx, void* dev, int64_t nbytes);
/* write `nbytes` of scratch larger than L2 (bench.py: flush L2 between timed iterations) */
int b2_flush_l2(b2_ctx* ctx);

/* ---- basis -----------------------------------------------------------------------------------
 * Tensor-product spline space on a structured topology.  Replaces function.StructuredBasis
 * (src/nutils/function.py:3040-3100: _coeffs, _start_dofs, _stop_dofs, _dofs_shape) as produced by
 * StructuredTopology.basis_spline (src/nutils/topology.py:2209-2324).  Per dimension d:
 *   coeffs[d]  float64[nsets[d]][degree[d]+1][degree[d]+1]  local polynomials, HIGHEST power first,
 *              in the element coordinate xi in [0,1]   (topology.py:2327-2361)
 *   setidx[d]  int32[nelems[d]]   coefficient set of each element
 *   start[d]   int64[nelems[d]]   first dof of each element; element dofs are start..start+degree
 *   ndofs[d]   number of dofs along d
 * Global dof = C-order ravel of the per-dimension dofs (function.py:3087-3093); vector-valued
 * fields with ncomp components use dof = ibasis*ncomp + comp (function.py:2623-2626). */
int b2_basis_create(b2_ctx* ctx, int ndims, const int64_t* nelems, const int32_t* degree, const int32_t* nsets,
                    const double* const* coeffs, const int32_t* const* setidx, const int64_t* const* start,
                    const int64_t* ndofs, int ncomp, b2_basis** out);
int b2_basis_destroy(b2_basis* basis);
int64_t b2_basis_ndofs(const b2_basis* basis);

/* ---- quadrature ------------------------------------------------------------------------------
 * Uniform tensor rule, the same for every element: replaces PointsSequence._Uniform over
 * TensorPoints of gauss1 rules (src/nutils/pointsseq.py:420-428, points.py:144-164, 342-355).
 * pts[d], wts[d]: float64[nq[d]] on [0,1]; tensor points are the C-order outer product. */
int b2_quad_create_tensor(b2_ctx* ctx, int ndims, const int32_t* nq, const double* const* pts, const double* const* wts, b2_quad** out);
int b2_quad_destroy(b2_quad* quad);

/* ---- geometry --------------------------------------------------------------------------------
 * Multilinear nodal geometry x = sum_v phi_v X_v, the geometry mesh.rectilinear builds for
 * non-integer vertices (src/nutils/mesh.py:55-57); J = dx/dxi, J^-1 and |det J| are evaluated per
 * quadrature point in the kernel (function.py:1207-1231, 1266-1295; evaluable.py:1403, 1463).
 * nodes_host: float64[ndims][nelems[0]+1][nelems[1]+1][nelems[2]+1]. */
int b2_geom_create_nodal(b2_ctx* ctx, int ndims, const int64_t* nelems, const double* nodes_host, b2_geom** out);
/* re-upload the coordinates (same shape); part of the end-to-end timed region in bench.py */
int b2_geom_update_nodal(b2_geom* geom, const double* nodes_host);
int b2_geom_destroy(b2_geom* geom);

/* ---- CSR pattern -----------------------------------------------------------------------------
 * The sorted, unique (row, col) set that the reference obtains after the element loop by
 * argsort/unique/compress_indices (src/nutils/evaluable.py:588-616, 5646-5682;
 * numeric.py:687-711) -- here built once, analytically, from the tensor structure of the basis;
 * structural zeros are kept exactly like the reference does. */
int b2_pattern_create(b2_ctx* ctx, const b2_basis* basis, b2_pattern** out);
int b2_pattern_destroy(b2_pattern* pattern);
int64_t b2_pattern_nnz(const b2_pattern* pattern);
int64_t b2_pattern_nrows(const b2_pattern* pattern);
/* offset of the first stored entry of `row` (0 <= row <= nrows), i.e. rowptr[row], computed analytically on the
 * host; lets a rank of a multi-GPU run allocate only its window of rows (see b2_assemble_device) */
int64_t b2_pattern_row_offset(const b2_pattern* pattern, int64_t row);
/* rowptr int64[nrows+1], colidx int64[nnz]; bit-equal to function.as_csr of the reference */
int b2_pattern_export_host(b2_pattern* pattern, int64_t* rowptr_host, int64_t* colidx_host);
int b2_pattern_export_device(b2_pattern* pattern, int64_t* rowptr_dev, int64_t* colidx_dev);

/* ---- assembly --------------------------------------------------------------------------------
 * One pass of the element loop that evaluable.compile generates (src/nutils/evaluable.py:6532-6838,
 * listing in SURVEY.md appendix A) for integrals lowered by sample._Integral.lower
 * (src/nutils/sample.py:944-956): basis evaluation (Polyval/PolyGrad), geometry, the integrand
 * einsums, the weighted sum over quadrature points and the scatter into CSR values
 * (LoopConcatenate + Assemble, evaluable.py:5383-5501, 3552-3620) and load vectors.
 *
 * Matrix form m is the bilinear form
 *     A[(i,c),(j,e)] = int sum_{x,y} D_m[c][x][e][y] d_x N_i d_y N_j |det J| dxi,   x,y in 0..ndims,
 * d_0 = value, d_k = derivative to physical coordinate k-1;  D_m: float64[ncomp][ndims+1][ncomp][ndims+1].
 * (mass: D[0][0][0][0]=1;  Laplace stiffness: D[0][k][0][k]=1, k>=1;  elasticity:
 *  D[c][1+k][e][1+l] = lambda d_ck d_el + mu (d_ce d_kl + d_cl d_ek).)
 * Vector form v:  b[(i,c)] = int sum_x C_v[c][x] d_x N_i |det J| dxi;  C_v: float64[ncomp][ndims+1].
 *
 * Elements [elem_begin, elem_end) of the C-order element numbering are integrated
 * (transformseq.py:563-579); pass 0, -1 for all.  values[m] (float64[nnz]) and rhs[v]
 * (float64[ndofs]) are ACCUMULATED INTO (zero them first, b2_memset_zero); this is what lets
 * ranks of a multi-GPU run integrate element slabs into their own arrays.  A rank that stores only the
 * rows [r0, r1) its slab touches passes values_dev[m] = window - b2_pattern_row_offset(pattern, r0) and
 * rhs_dev[v] = window - r0 (in elements): only addresses inside the window are accessed. */
int b2_assemble_device(b2_ctx* ctx, const b2_pattern* pattern, const b2_basis* basis, const b2_quad* quad, const b2_geom* geom,
                       int64_t elem_begin, int64_t elem_end,
                       int nmat, const double* const* D_host, double* const* values_dev,
                       int nvec, const double* const* C_host, double* const* rhs_dev);
/* Owner-computes variant of the same integrals: every stored CSR value (and load-vector entry) of the dof rows whose
 * index along dimension 0 lies in [plane_begin, plane_end) is WRITTEN exactly once -- the sum over the elements in
 * the common support of N_i and N_j, which the reference forms after its loop by argsort/unique/accumulate
 * (src/nutils/evaluable.py:588-616, 5646-5682; numeric.py:434-460), is folded into the integration itself.  No
 * zero-fill is needed and rows outside the planes are not touched, so the ranks of a multi-GPU run that own disjoint
 * plane ranges need NO exchange step (each integrates the <= degree element layers below its first plane again).
 * Pass 0, -1 for all planes.  values_dev / rhs_dev follow the windowing convention of b2_assemble_device. */
int b2_assemble_rows_device(b2_ctx* ctx, const b2_pattern* pattern, const b2_basis* basis, const b2_quad* quad, const b2_geom* geom,
                            int64_t plane_begin, int64_t plane_end,
                            int nmat, const double* const* D_host, double* const* values_dev,
                            int nvec, const double* const* C_host, double* const* rhs_dev);
/* Host-buffer variant: zero-fills device scratch owned by the context, assembles, copies the
 * results to the caller's host buffers (pinned memory from b2_host_alloc recommended). */
int b2_assemble_host(b2_ctx* ctx, const b2_pattern* pattern, const b2_basis* basis, const b2_quad* quad, const b2_geom* geom,
                     int64_t elem_begin, int64_t elem_end,
                     int nmat, const double* const* D_host, double* const* values_host,
                     int nvec, const double* const* C_host, double* const* rhs_host);
/* ---- element sets: trimmed topologies, ragged quadrature, pruned and rational bases ----------------
 * The structured entry points above integrate EVERY element of the grid with ONE tensor rule.  An element set
 * describes what the reference builds for the other topologies that still live on a structured grid:
 *   - a subset of the elements (SubsetTopology / trimmed topologies, src/nutils/topology.py:1598-1660; the element
 *     indices are what PrunedBasis keeps as _transmap, function.py:3118-3123);
 *   - per-element point sets of different length, in element-local coordinates (PointsSequence._Plain over
 *     ConcatPoints of the cut cells, pointsseq.py:324-332, points.py:257-337, element.py:669-680, 861-873);
 *   - the pruned dof numbering (PrunedBasis._renumber, function.py:3121-3133): renumber[parent basis function] = new
 *     index, or -1 (or >= nbasis_new) for functions without support on the subset; must be increasing over the kept
 *     functions, which is what numeric.invmap of a sorted dof list gives;
 *   - rational (NURBS) functions R_i = c_i B_i / W (examples/platewithhole.py:66-83): `scale` holds c_i per PARENT
 *     basis function; `rational` selects the denominator: 0 none (scale only), 1 W = sum_j c_j B_j (the weight function
 *     in the solution basis), 2 the weight function of a rational spline geometry (b2_geom_create_spline).
 * elem_ids: int64[nsel], strictly increasing C-order element indices (NULL = all elements, nsel = their number).
 * qoff: int64[nsel+1] offsets into qcoords (float64[npoints][ndims], xi in [0,1]) and qweights (float64[npoints]);
 * NULL = every element uses the tensor rule passed at assembly time.  All arrays are copied. */
typedef struct b2_elemset b2_elemset;
int b2_elemset_create(b2_ctx* ctx, const b2_basis* basis, int64_t nsel, const int64_t* elem_ids,
                      const int64_t* qoff, const double* qcoords, const double* qweights,
                      const int64_t* renumber, int64_t nbasis_new, const double* scale, int rational, b2_elemset** out);
int b2_elemset_destroy(b2_elemset* elemset);
int64_t b2_elemset_ndofs(const b2_elemset* elemset);   /* nbasis_new * ncomp */
int64_t b2_elemset_npoints(const b2_elemset* elemset); /* total number of quadrature points (0 for tensor rules) */

/* Boundary integrals and pointwise coefficients (Neumann terms and the boundary projections of
 * solve_constraints integrate over boundary topologies, src/nutils/topology.py:2049-2057, sample.py:944-956 on a
 * (d-1)-dimensional sample; coefficient functions are evaluated at the points, function.py:2291-2316 for the measure):
 *   b2_elemset_set_faces: face_dim[k] (int8[nsel]) = -1 for a volume element set entry, or the reference direction normal
 *     to the face the entry's points lie on (their local coordinate along that direction is 0 or 1).  The weight of such a
 *     point is multiplied by the SURFACE measure |det J| |J^-T e_dim| instead of |det J| (function.J on a boundary sample).
 *   b2_elemset_set_coefficient: a scalar per point that multiplies one form of the next assembly calls: which = m for
 *     matrix form m, B2_MAX_FORMS + v for vector form v; coef float64[npoints] in point order (element sets with a tensor
 *     rule: [nsel][nq]), NULL removes it.  The array is copied.
 *   b2_elemset_set_normals: the general form of set_faces for facets that are not aligned with the element faces -- the
 *     immersed boundary of a trimmed topology (topo.boundary['trimmed'], the simplices of the cut-cell mosaics).  nref
 *     (float64[npoints][ndims]) is, per point, the reference-space normal of its facet SCALED by the facet's reference measure
 *     (cross product of the columns of the facet-to-element linear map, transform.py); the weight of the point is multiplied by
 *     |det J| |J^-T nref| (Nanson).  NULL removes it; it takes precedence over face_dim. */
int b2_elemset_set_faces(b2_elemset* elemset, const int8_t* face_dim);
int b2_elemset_set_normals(b2_elemset* elemset, const double* nref, int64_t npoints);
int b2_elemset_set_coefficient(b2_elemset* elemset, int which, const double* coef, int64_t npoints);

/* Spline geometry x(xi) = sum_i B_i(xi) X_i, or the rational map sum_i B_i w_i X_i / sum_i B_i w_i (the NURBS map of
 * examples/platewithhole.py:66-78 represented in the basis of the refined topology): gbasis is a scalar b2_basis on
 * the same element grid, ctrl_host float64[ndims][nbasis] the control points X_i, weights_host float64[nbasis] or NULL.
 * J = dx/dxi follows from the derivatives of the basis (function.py:1207-1231 applied to a spline geometry). */
int b2_geom_create_spline(b2_ctx* ctx, const b2_basis* gbasis, const double* ctrl_host, const double* weights_host, b2_geom** out);

/* CSR pattern of an element set: the sorted unique (row, col) pairs of the dofs that share a SELECTED element -- what
 * the reference finds by argsort/unique over the COO keys (evaluable.py:588-616, 5646-5682) -- built on the device
 * from the tensor structure (candidate box per row, kept if a selected element lies in the common support), two
 * passes (count, scan, fill).  rowptr/colidx are materialised on the device; b2_pattern_export_*, _nnz, _nrows and
 * _row_offset work on it like on the analytic pattern. */
int b2_pattern_create_elemset(b2_ctx* ctx, const b2_elemset* elemset, b2_pattern** out);

/* Solution-dependent coefficients (what System.assemble_jacobian_residual re-evaluates per Newton step,
 * src/nutils/solver.py:357-425): matrix form `which` (or vector form which - B2_MAX_FORMS) is multiplied at every point by
 * scale * u_h(x_q)^power, with u_h = sum_i field_dev[i] N_i evaluated INSIDE the assembly kernel from the caller-owned device
 * vector field_dev (float64[ndofs], read at launch time: update it in place between Newton steps).  Combines with
 * b2_elemset_set_coefficient (product).  field_dev NULL removes it.  Scalar, non-rational spaces. */
int b2_elemset_set_coefficient_field(b2_elemset* elemset, int which, const double* field_dev, int power, double scale);

/* Element-scatter assembly of the selected elements [sel_begin, sel_end) (positions in elem_ids; 0, -1 = all):
 * same forms and the same ACCUMULATE semantics as b2_assemble_device.  quad may be NULL when the element set
 * carries its own points.  The slot of (row, col) is found by bisection in the row's sorted column list. */
int b2_assemble_elemset_device(b2_ctx* ctx, const b2_pattern* pattern, const b2_elemset* elemset, const b2_quad* quad, const b2_geom* geom,
                               int64_t sel_begin, int64_t sel_end,
                               int nmat, const double* const* D_host, double* const* values_dev,
                               int nvec, const double* const* C_host, double* const* rhs_dev);
/* zero-fill + b2_assemble_elemset_device + copy to host buffers */
int b2_assemble_elemset_host(b2_ctx* ctx, const b2_pattern* pattern, const b2_elemset* elemset, const b2_quad* quad, const b2_geom* geom,
                             int nmat, const double* const* D_host, double* const* values_host,
                             int nvec, const double* const* C_host, double* const* rhs_host);

/* Sample.eval / Sample.bind (src/nutils/sample.py:192-237, 959-975): evaluate at every point of the element set, in point
 * order (elements in selection order, points in their stored / C-order tensor order): the physical coordinates x
 * (float64[npoints][ndims]), the integration weight w |det J| -- surface measure on faces -- (float64[npoints]), and for
 * `nfields` coefficient vectors coef_dev (float64[nfields][ndofs]) the discrete fields u = sum_i coef[i] N_i
 * (values float64[npoints][nfields * ncomp]) and their physical gradients (float64[npoints][nfields * ncomp][ndims]).
 * Any output pointer may be NULL.  This is what error norms, plots and pointwise post-processing of the reference's
 * examples consume; with it an integral of ANY pointwise function of the solution is a dot product on the host. */
int b2_evaluate_elemset_device(b2_ctx* ctx, const b2_elemset* elemset, const b2_quad* quad, const b2_geom* geom, int nfields, const double* coef_dev,
                               double* x_dev, double* wdet_dev, double* values_dev, double* grads_dev);

/* ---- device-resident matrix operations -------------------------------------------------------------
 * The assembled values stay in HBM (a 128^3 p=2 matrix is 2.1 GB: copying it to the host costs 10x the assembly);
 * these entry points are what nutils.matrix.Matrix offers on the result (src/nutils/matrix/_base.py), for BOTH kinds
 * of pattern, without ever materialising colidx for the analytic one:
 *   b2_spmv_device      y = A x                      Matrix.__matmul__            (_base.py:62-75)
 *   b2_diagonal_device  diag(A)                      Matrix.diagonal              (_base.py:86-90)
 *   b2_cg_device        Jacobi-preconditioned CG     Matrix.solve(rhs, lhs0=..., constrain=..., solver='cg', precon='diag',
 *                       atol=..., rtol=...)          (_base.py:100-173, 199-213; used by solver.System.solve, solver.py:318-425)
 * b2_cg_device: x_dev holds lhs0 on entry -- INCLUDING the values of the constrained dofs -- and the solution on return;
 * constrained_dev (uint8[nrows], may be NULL) marks constrained dofs (rows and columns), rhs_dev may be NULL (zero).
 * Solves A_ff dx = (b - A lhs0)_f and stops when |r| <= max(atol, rtol |(b - A lhs0)_f|); atol = rtol = 0 means "to
 * machine precision" (1e-13 relative) like the reference.  maxiter 0 = automatic.  Returns B2_ENOTCONVERGED (x holds the
 * best iterate, the reference's ToleranceNotReached.best) if an explicit tolerance was not reached. */
int b2_spmv_device(b2_ctx* ctx, const b2_pattern* pattern, const double* values_dev, const double* x_dev, double* y_dev);
int b2_diagonal_device(b2_ctx* ctx, const b2_pattern* pattern, const double* values_dev, double* diag_dev);
int b2_cg_device(b2_ctx* ctx, const b2_pattern* pattern, const double* values_dev, const double* rhs_dev, double* x_dev,
                 const unsigned char* constrained_dev, double atol, double rtol, int maxiter, int* iterations, double* resnorm);

/* ---- general dof maps ---------------------------------------------------------------
 * b2_pattern_create_csr: a pattern from explicit CSR index arrays (int64, rowptr[0] = 0, columns strictly increasing per
 * row and < ncols): what nutils.matrix.assemble_csr(values, rowptr, colidx, ncols) validates and hands to a backend
 * (src/nutils/matrix/__init__.py:30-70).  b2_spmv_device / b2_diagonal_device / b2_cg_device / b2_pattern_export_* accept it.
 *
 * b2_pattern_general: the pattern of  sum_e rowdofs(e) x coldofs(e)  for explicit element dof lists (CSR-of-lists:
 * rowdofs[rowoff[e] .. rowoff[e+1])), any basis -- simplex meshes (topology.py:2493 SimplexTopology.basis_std), mixed
 * meshes (mesh.py:737-753), hierarchical bases.  Replaces the reference's post-loop flatten / stable argsort / unique /
 * compress_indices (evaluable.py:588-616, 5646-5682; numeric.py:687-711): keys row << 32 | col are generated, radix-sorted
 * and made unique on the device.  dof lists hold BASIS indices; with ncomp > 1 the pattern is expanded with the field
 * numbering dof = basis * ncomp + component (function.py:2623-2626).  rowptr/colidx (b2_pattern_export_*) are bit-equal to
 * evaluable.as_csr. */
int b2_pattern_create_csr(b2_ctx* ctx, int64_t nrows, int64_t ncols, const int64_t* rowptr, const int64_t* colidx, b2_pattern** out);
int b2_pattern_general(b2_ctx* ctx, int64_t nelems, const int64_t* rowoff, const int64_t* rowdofs, const int64_t* coloff, const int64_t* coldofs,
                       int64_t nrows, int64_t ncols, int ncomp, b2_pattern** out);

/* Element loop on a general pattern for element types described by TABULATED reference data -- the generated loop of the
 * reference (evaluable.py:6773-6787) for bases given as per-element coefficient arrays (function.PlainBasis,
 * function.py:2881-2913).  Per element type t (etype[e] selects it): nq[t] points with weights[t][q]; the nfun[t] local
 * functions and their reference gradients phi[t][q][a], dphi[t][q][a][k]; the nvert[t] geometry shape functions'
 * reference gradients gdphi[t][q][v][k] (linear on simplices, multilinear on tensor cells; gphi is accepted for symmetry
 * and not needed).  Per element: its dofs (basis indices, CSR-of-lists dofoff/dofs, the lists given to
 * b2_pattern_general) and its vertex coordinates vertcoords[vertoff[e] + v][ndims].  Forms D, C as in b2_assemble_device.
 * Zero-fills, integrates every element (J, J^-1, |det J| per point), scatters with fp64 atomics and copies to the host. */
int b2_assemble_general_host(b2_ctx* ctx, const b2_pattern* pattern, int ndims, int ncomp, int64_t nelems, int ntypes, const int32_t* etype,
                             const int32_t* nq, const int32_t* nfun, const int32_t* nvert, const double* const* weights, const double* const* phi,
                             const double* const* dphi, const double* const* gphi, const double* const* gdphi, const int64_t* dofoff, const int64_t* dofs,
                             const int64_t* vertoff, const double* vertcoords, int nmat, const double* const* D_host, double* const* values_host, int nvec,
                             const double* const* C_host, double* const* rhs_host);

/* Experiments and profiling.  "kernel": 0 = automatic, 1 = generic (coverage) kernel only, 2 = specialised kernel or
 * B2_EUNSUPPORTED.  "path": 0 = b2_assemble_host uses the owner-computes rows path for the whole topology, 1 = always
 * the element-scatter path.  "rows_nseg": force the number of marching segments of the rows kernel (0 = automatic).
 * "time_kernels": see b2_ctx_kernel_time (2: report the longest launch instead of the sum).  "rows_gpre": 0 = the rows
 * kernel evaluates the geometry itself instead of fetching the precomputed array by TMA.  "rows_sym": 1 / 0 = scalar forms
 * on unordered / ordered dof pairs (default: per degree).  "rows_vecsym": 0 = symmetric vector-valued forms integrate all
 * nine component blocks instead of the six on and above the diagonal.  "elemset_mma": 0 = scalar FMA block loop.
 * "elemset_queue": 0 = elements of ragged sets dealt round robin to the CTAs instead of the dynamic longest-first queue. */
int b2_ctx_set_option(b2_ctx* ctx, const char* name, int64_t value);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* B200FEM_H */
