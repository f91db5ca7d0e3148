/* msthost.h -- C ABI of libmsthost.so: the host-side companions of the GPU path (no CUDA).
 *
 * Everything a host needs on either side of the solver when it does not carry the reference's
 * pointer-graph mesh (R = /root/reference/MST-CFD).  What each entry point replaces:
 *
 *   msthost_msh_read / _parse / _sizes / _tables / _zone_name / _free
 *                         the reader half of MshBlock::readMsh (R/mesh/MshBlock.cpp:75-271):
 *                         Fluent ASCII .msh subset -> raw tables (nodes, face -> nodes, c0, c1, zones)
 *   msthost_flatten       the metrics half (R/mesh/Face.cpp:8-44,62-69, R/mesh/Cell.cpp:6-61,
 *                         R/mesh/MshBlock.cpp:281-334): raw tables -> the arrays of mstgpu_mesh
 *   msthost_node_faces    Node::addNbFace order (R/mesh/Node.cpp:13-15)        -> mstgpu_output_setup
 *   msthost_cell_nodes    Cell::getBeginItPNbNodes order (MshBlock.cpp:335-368) -> the element list
 *   msthost_plt_write     the file Work::writedataRhoBasedMshNodePlt writes (R/work/Work.cpp:204-319),
 *                         byte for byte, from the node fields of mstgpu_node_fields
 *   msthost_plt_write_binary   same content, raw doubles
 *   msthost_msh_write     raw tables -> a .msh file the reference's own reader accepts
 *   msthost_box_tets / msthost_grid_tris (+ _sizes)   synthetic meshes of the BASELINE configs
 *
 * All arrays are caller-allocated host memory.  Return value: 0 = ok, negative = error;
 * msthost_last_error() gives the text of the calling thread's last failure.
 */
#ifndef MSTHOST_H
#define MSTHOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct msthost_msh msthost_msh; /* a parsed .msh file */

const char* msthost_last_error(void);

/* ---- mesh input ---------------------------------------------------------------------------------- */
int msthost_msh_read(const char* path, msthost_msh** out);
int msthost_msh_parse(const char* text, int64_t nbytes, msthost_msh** out); /* same on a memory image */
void msthost_msh_free(msthost_msh* h);
/* sizes8 = dim, nnodes, ncells, nfaces, nint (MshBlock::getNumOfIntFaces), nzones, npf (row stride of
 * face_nodes = most nodes per face), 0 */
int msthost_msh_sizes(const msthost_msh* h, int64_t* sizes8);
/* any pointer may be NULL.  nodes [nnodes*dim]; face_nodes [nfaces*npf] 0-based, -1 padded; c0, c1 [nfaces]
 * 0-based, c1 = -1 on boundary faces; ftype [nfaces] zone type of the face; zones [nzones*5] =
 * id, first face (0-based), end (exclusive), type, nodes per face */
int msthost_msh_tables(const msthost_msh* h, double* nodes, int32_t* face_nodes, int32_t* c0, int32_t* c1,
                       int32_t* ftype, int32_t* zones);
const char* msthost_msh_zone_name(const msthost_msh* h, int32_t zone); /* FacesInf::getName */

/* flag_convention: 0 = consistent (flag[d] = Sout_c0[d] >= 0), 1 = as the reader builds them
 * (MshBlock.cpp:284-303).  Outputs: S, fc [nfaces*dim]; dac, eta [nfaces]; flag [nfaces*dim];
 * cc [ncells*dim]; vol [ncells]; cf_ptr [ncells+1]; cf_idx [nfaces + interior faces]. */
int msthost_flatten(int dim, int64_t nnodes, int64_t ncells, int64_t nfaces, int npf, const double* nodes,
                    const int32_t* face_nodes, const int32_t* c0, const int32_t* c1, int flag_convention,
                    double* S, double* fc, int8_t* dac, double* eta, uint8_t* flag, double* cc, double* vol,
                    int32_t* cf_ptr, int32_t* cf_idx);

/* nf_ptr [nnodes+1], nf_idx [number of (face, node) pairs] */
int msthost_node_faces(int64_t nnodes, int64_t nfaces, int32_t npf, const int32_t* face_nodes, int32_t* nf_ptr,
                       int32_t* nf_idx);

/* ---- output -------------------------------------------------------------------------------------- */
/* cn_ptr [ncells+1] is always filled; cn_idx (capacity cn_ptr[ncells]) may be NULL on a sizing call */
int msthost_cell_nodes(int32_t dim, int64_t ncells, int32_t npf, const int32_t* face_nodes, const int32_t* cf_ptr,
                       const int32_t* cf_idx, const double* nodes, int32_t* cn_ptr, int32_t* cn_idx);
/* fields [nnodes][dim+4] = rho, u_i, T, p, Ma; zone_t = the step counter printed in ZONE T="...";
 * felnum = FELNUM of CONST.h:5 (3 -> FETRIANGLE, else FEQUADRILATERAL; 3-D: FETETRAHEDRON) */
int msthost_plt_write(const char* path, int32_t dim, int64_t nnodes, int64_t ncells, const double* nodes,
                      const double* fields, const int32_t* cn_ptr, const int32_t* cn_idx, int32_t zone_t,
                      int32_t felnum);
int msthost_plt_write_binary(const char* path, int32_t dim, int64_t nnodes, int64_t ncells, const double* nodes,
                             const double* fields, const int32_t* cn_ptr, const int32_t* cn_idx, int32_t zone_t);
/* zones [nzones*5] as in msthost_msh_tables */
int msthost_msh_write(const char* path, int32_t dim, int64_t nnodes, int64_t ncells, int64_t nfaces, int32_t npf,
                      const double* nodes, const int32_t* face_nodes, const int32_t* c0, const int32_t* c1,
                      int32_t nzones, const int32_t* zones);

/* ---- synthetic meshes (BASELINE configs 2, 4, 5) --------------------------------------------------- */
void msthost_box_tets_sizes(int nx, int ny, int nz, int64_t* nnodes, int64_t* ncells, int64_t* nfaces,
                            int64_t* nint);
/* bc[6] = zone types of the sides x-, x+, y-, y+, z-, z+ */
int msthost_box_tets(int nx, int ny, int nz, double lx, double ly, double lz, const int32_t* bc, double* nodes,
                     int32_t* face_nodes, int32_t* c0, int32_t* c1, int32_t* ftype);
void msthost_grid_tris_sizes(int nx, int ny, const uint8_t* mask, int64_t* nnodes, int64_t* ncells,
                             int64_t* nfaces, int64_t* nint);
/* bc[3] = zone type at x = 0, at x = lx, elsewhere */
int msthost_grid_tris(int nx, int ny, double lx, double ly, const uint8_t* mask, const int32_t* bc, double* nodes,
                      int32_t* face_nodes, int32_t* c0, int32_t* c1, int32_t* ftype);

#ifdef __cplusplus
}
#endif
#endif /* MSTHOST_H */
