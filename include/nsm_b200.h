/* include/nsm_b200.h — C ABI of the B200-native hex8 explicit-dynamics path.
 *
 * This is the drop-in boundary: plain C types, caller-owned host buffers, no torch /
 * Kokkos types.  The C++ host classes in nimblesm_b200/host (B200ModelData, B200Block,
 * B200BlockMaterialInterface, ...) and the ctypes mirror nimblesm_b200/capi.py call exactly these
 * entry points.  Every entry point names the reference interface (path:line under /root/reference)
 * that it replaces.  All work is issued on one CUDA stream owned by the context; calls are
 * synchronous unless stated.  There is no CPU fallback: every call fails with NSM_ERR_CUDA when no
 * sm_100 device is usable.
 *
 * Conventions
 *   - status: 0 = NSM_OK, otherwise an nsm_status; nsm_b200_last_error() gives the message.
 *     Nothing throws across this boundary (reference: NIMBLE_ABORT / std::invalid_argument,
 *     src/nimble_macros.h:49-72, src/nimble.cc:129-137 — the C++ wrappers convert).
 *   - host nodal VECTOR fields are AoS [n_nodes][3] doubles exactly as nimble::Viewify<2> with
 *     strides {3,1} presents them (src/nimble_model_data.cc:540-546); SCALAR fields are [n_nodes].
 *     On the device everything is SoA fp64 (x[], y[], z[] separately).
 *   - connectivity is int32 [n_elem][8], 0-based local node ids, Exodus hex8 ordering
 *     (src/nimble_element.cc:113-120).
 *   - integration-point data is [n_elem][8][15]: F in storage order xx,yy,zz,xy,yz,zx,yx,zy,xz then
 *     sigma xx,yy,zz,xy,yz,zx (src/nimble_block.cc:84-108, src/nimble_utils.h:86-110).
 */
#ifndef NSM_B200_H
#define NSM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nsm_b200_ctx nsm_b200_ctx;

typedef enum {
  NSM_OK          = 0,
  NSM_ERR_ARG     = 1, /* bad argument / call order                                       */
  NSM_ERR_CUDA    = 2, /* CUDA runtime error or no usable sm_100 device                   */
  NSM_ERR_JACOBIAN = 3, /* non-positive Jacobian determinant met (src/nimble_utils.h:1253) */
  NSM_ERR_MATERIAL = 4, /* unknown material model / parameter                              */
  NSM_ERR_COMM    = 5  /* peer exchange not set up / peer failure                         */
} nsm_status;

/* Material models.  ELASTIC and NEOHOOKEAN are the reference's (src/nimble_material.cc:60,218; both carry 0 state
 * variables).  J2_PLASTICITY fills the reference's state-variable slot (Material::NumStateVariables /
 * GetStateVariableLabel / GetStateVariableInitialValue, src/nimble_material.h:214-225; N / N+1 element data,
 * src/nimble_block.cc:297-368): small-strain J2 plasticity with linear isotropic hardening in incremental form,
 * sigma_np1 = radial_return(sigma_n + C : (sym F_np1 - sym F_n)), two state variables per integration point
 * (equivalent_plastic_strain, von_mises_stress).  Its operation sequence is pinned by a nimble::Material subclass
 * that the checker plugs into the reference's unmodified block code (oracle/ref_state_material.cc). */
typedef enum { NSM_MAT_ELASTIC = 0, NSM_MAT_NEOHOOKEAN = 1, NSM_MAT_J2_PLASTICITY = 2, NSM_MAT_COUNT = 3 } nsm_material_kind;

/* nodal fields allocated by DataManager::Initialize (src/nimble_data_manager.cc:135-158). */
typedef enum {
  NSM_FIELD_LUMPED_MASS          = 0, /* scalar */
  NSM_FIELD_REFERENCE_COORDINATE = 1,
  NSM_FIELD_DISPLACEMENT         = 2,
  NSM_FIELD_VELOCITY             = 3,
  NSM_FIELD_ACCELERATION         = 4,
  NSM_FIELD_INTERNAL_FORCE       = 5,
  NSM_FIELD_EXTERNAL_FORCE       = 6,
  NSM_FIELD_CONTACT_FORCE        = 7, /* zero until nsm_b200_set_contact; download only */
  NSM_FIELD_COUNT                = 8
} nsm_field;

/* boundary-condition kinds applied inside the step (src/nimble_boundary_condition.h:62-69). */
typedef enum { NSM_BC_PRESCRIBED_VELOCITY = 0, NSM_BC_PRESCRIBED_DISPLACEMENT = 1 } nsm_bc_kind;

/* force assembly (SURVEY.md §2b "scatter-add"):
 *   ATOMIC  : red.global.add.f64 per nodal component, order of the <=8 element contributions per node
 *             is not fixed (run-to-run noise ~1e-16 relative).
 *   ORDERED : element kernel stores per-element nodal forces, the node kernel sums them through a
 *             node->element adjacency in ascending (block id, element) order == the summation order
 *             of the reference's serial loop (src/nimble_model_data.cc:636-659, src/nimble_block.cc:434),
 *             so nodal forces are bit-reproducible and bit-identical to the serial reference. */
typedef enum { NSM_ASSEMBLY_ATOMIC = 0, NSM_ASSEMBLY_ORDERED = 1 } nsm_assembly;

/* nsm_b200_finalize flags */
#define NSM_FLAG_STORE_IPT_EVERY_STEP 0x1 /* write F/sigma each step as the reference does (960 B/elem);     \
                                             default: only when a call asks for it (output steps)          */
#define NSM_FLAG_CACHE_REF_JACOBIAN 0x2   /* keep inverse reference Jacobians (576 B/elem) instead of      \
                                             recomputing them each step; bit-identical either way          */
#define NSM_FLAG_REORDER_ELEMENTS 0x4     /* walk each block's elements along a Morton curve of their       \
                                             centroids instead of in file order (internal schedule only:    \
                                             element data, outputs and the ORDERED summation keep the file  \
                                             order).  For meshes whose element order has no locality: a     \
                                             randomly ordered 8 M-element cube runs 3.4x slower without it. */
#define NSM_FLAG_RENUMBER_NODES 0x8       /* number the nodes along a Morton curve of their coordinates     \
                                             INSIDE the context; every entry point keeps speaking the        \
                                             caller's node ids (fields, BC table, shared-node lists are      \
                                             mapped at the boundary).  For meshes whose node numbering has   \
                                             no locality; use together with NSM_FLAG_REORDER_ELEMENTS.       */

/* ---- lifetime ------------------------------------------------------------------------------- */
/* Creates a context bound to CUDA device `device` (one context per GPU, one host thread drives it:
 * the reference is single-threaded per rank, src/nimble.cc:139-143). */
int nsm_b200_create(int device, nsm_b200_ctx** out);
void nsm_b200_destroy(nsm_b200_ctx* ctx);
/* Message of the last failure on ctx (ctx may be NULL: creation failures). */
const char* nsm_b200_last_error(const nsm_b200_ctx* ctx);
/* Library build info: "nsm_b200 <version> sm_100a fmad=off ...". */
const char* nsm_b200_version(void);

/* ---- model definition (replaces ModelData::SetReferenceCoordinates + InitializeBlocks/EmplaceBlocks,
 *      src/nimble_model_data.cc:416-493, src/nimble_kokkos_model_data.cc:893-961) ------------------ */
int nsm_b200_set_nodes(nsm_b200_ctx* ctx, int64_t n_nodes, const double* x, const double* y, const double* z);
/* Blocks may be added in any order; they are processed in ascending block_id like the reference's
 * std::map (src/nimble_model_data.cc:636).  Material = MaterialFactory::create result
 * (src/nimble_material_factory.cc:56-69): bulk_modulus, shear_modulus, density. */
int nsm_b200_add_block(nsm_b200_ctx* ctx, int block_id, int64_t n_elem, const int32_t* conn, int material_kind,
                       double bulk_modulus, double shear_modulus, double density);
/* The same with the material's full parameter list: params = {bulk_modulus, shear_modulus, density} followed by the
 * model's own parameters (J2_PLASTICITY: yield_stress, hardening_modulus); n_params must equal
 * nsm_b200_material_num_params(kind).  A block whose material has state variables keeps two integration-point
 * record arrays [n_elem][8][15 + n_state] on the device (N and N+1, Block::InitializeElementData gives both the
 * initial values, src/nimble_block.cc:148-207) and writes F / sigma / state every step. */
int nsm_b200_add_block_params(nsm_b200_ctx* ctx, int block_id, int64_t n_elem, const int32_t* conn, int material_kind,
                              int n_params, const double* params);
/* Material::NumStateVariables / GetStateVariableLabel / GetStateVariableInitialValue (src/nimble_material.h:214-225)
 * and the parameter count of a kind (-1 / NULL for an unknown kind or index). */
int         nsm_b200_material_num_state(int material_kind);
int         nsm_b200_material_num_params(int material_kind);
const char* nsm_b200_material_state_label(int material_kind, int index);
double      nsm_b200_material_state_initial_value(int material_kind, int index);
/* Uploads the mesh, builds assembly tables, allocates fields (zeroed; F = identity, sigma = 0 like
 * Block::InitializeElementData, src/nimble_block.cc:148-207). */
int nsm_b200_finalize(nsm_b200_ctx* ctx, int assembly, unsigned flags);

int64_t nsm_b200_num_nodes(const nsm_b200_ctx* ctx);
int64_t nsm_b200_num_elements(const nsm_b200_ctx* ctx, int block_id); /* block_id < 0: all blocks */
int64_t nsm_b200_device_bytes(const nsm_b200_ctx* ctx);               /* HBM held by the context  */
/* The flags in force after nsm_b200_finalize: NSM_FLAG_CACHE_REF_JACOBIAN is dropped (the reference Jacobians are
 * recomputed every step, same bits) when the 576 B/element cache would not leave room for the integration-point
 * records of an output step beside what is already resident. */
unsigned nsm_b200_effective_flags(const nsm_b200_ctx* ctx);

/* ---- nodal fields (replaces ModelData::GetNodeData / UpdateWithNewVelocity / UpdateWithNewDisplacement
 *      host<->device deep copies, src/nimble_kokkos_model_data.cc:1730-1740,1282) ------------------ */
int nsm_b200_upload_field(nsm_b200_ctx* ctx, int field, const double* host);
int nsm_b200_download_field(nsm_b200_ctx* ctx, int field, double* host);
/* Asynchronous variants on the context stream (host memory should be pinned); pair with nsm_b200_sync. */
int nsm_b200_upload_field_async(nsm_b200_ctx* ctx, int field, const double* host);
int nsm_b200_download_field_async(nsm_b200_ctx* ctx, int field, double* host);
int nsm_b200_sync(nsm_b200_ctx* ctx);
/* Pinned host allocation helpers for callers without a CUDA runtime of their own. */
void* nsm_b200_host_alloc(int64_t bytes);
void  nsm_b200_host_free(void* p);

/* ---- setup kernels ---------------------------------------------------------------------------- */
/* ModelData::ComputeLumpedMass (src/nimble_model_data.cc:495-530 -> src/nimble_block.cc:110-146 ->
 * src/nimble_element.h:266-309) and BlockBase::ComputeCriticalTimeStep (src/nimble_block_base.cc:51-84,
 * min over blocks/elements; informational).  Includes the shared-node sum when peers are attached. */
int nsm_b200_compute_lumped_mass(nsm_b200_ctx* ctx, double* critical_dt);

/* ---- internal force (replaces ModelDataBase::ComputeInternalForce, src/nimble_model_data.cc:620-667,
 *      src/nimble_kokkos_model_data.cc:1154-1286: gather -> F -> stress -> Bt.sigma.detJ.w -> scatter ->
 *      shared-node sum) --------------------------------------------------------------------------- */
/* From the device-resident displacement into the device-resident internal_force. store_ipt != 0 also
 * writes F/sigma of every integration point (is_output_step of the reference). */
int nsm_b200_internal_force(nsm_b200_ctx* ctx, int store_ipt);
/* Same through host views: uploads displacement [n][3], computes, downloads internal_force [n][3]
 * (the exact shape of the reference call with Viewify<2> arguments, src/nimble_model_data_base.h:208-215; the
 * reference's Kokkos path crosses here every step: deep_copy of displacement and velocity up, internal force down,
 * src/nimble_kokkos_model_data.cc:1730-1740, 1282).  PIPELINED over the node chunks of nsm_b200_step_host
 * (nsm_b200_set_host_step_chunks): chunk k of the displacement travels up while the elements below it run and the
 * force of every node chunk whose elements have all run travels down; same kernels, same bits as the plain schedule.
 * Contexts with a peer exchange or an internal node renumbering take the plain schedule. */
int nsm_b200_internal_force_host(nsm_b200_ctx* ctx, const double* displacement, double* internal_force,
                                 int store_ipt);

/* ---- stress seam (replaces BlockMaterialInterface::ComputeStress, the MDRange(elem, ipt) loop of
 *      src/nimble_kokkos_block_material_interface.cc:65-119 -> Material::GetStress,
 *      src/nimble_material.cc:95-126, 252-310) ----------------------------------------------------- */
/* Host arrays: def_grad [n_points][9] -> stress [n_points][6], evaluated on the device.  PARITY / PLUG-IN SEAM: the
 * reference re-creates and calls this seam every step (src/nimble_kokkos_model_data.cc:1230-1234); here the step itself
 * never leaves the device, and a caller who does cross the seam pays two copies and a launch per call (the device
 * scratch is kept and only grows). */
int nsm_b200_compute_stress(nsm_b200_ctx* ctx, int material_kind, double bulk_modulus, double shear_modulus,
                            int64_t n_points, const double* def_grad, double* stress);

/* The full seam for a material with state variables: (F_n, F_np1, sigma_n, state_n) -> (sigma_np1, state_np1), host
 * arrays [n_points][9], [n_points][6], [n_points][n_state] (the four views compute_block_stress hands to
 * Material::GetStress, src/nimble_kokkos_block_material_interface.cc:86-118).  Materials without state ignore the N
 * inputs.  PARITY / PLUG-IN SEAM, not a performance path: it allocates, copies and frees per call; the step keeps
 * everything on the device. */
int nsm_b200_compute_stress_state(nsm_b200_ctx* ctx, int material_kind, int n_params, const double* params, int64_t n_points,
                                  const double* def_grad_n, const double* def_grad_np1, const double* stress_n,
                                  const double* state_n, double* stress_np1, double* state_np1);

/* ---- boundary conditions (replaces BoundaryConditionManager::ApplyKinematicBC,
 *      src/nimble_boundary_condition_manager.h:136-204) -------------------------------------------- */
/* Flattened table in deck order: entry k constrains velocity(node[k], comp[k]).  When a (node, comp)
 * pair appears more than once the later entry wins, as in the reference's sequential loop.
 * value[k] = prescribed velocity, or prescribed displacement d for which v = (d - u)/dt when dt > 0. */
int nsm_b200_set_bc_table(nsm_b200_ctx* ctx, int64_t n, const int32_t* node, const int32_t* comp,
                          const int32_t* kind);
/* Host-evaluated magnitudes (constants or expression(x,y,z,t) results) for the NEXT steps. */
int nsm_b200_set_bc_values(nsm_b200_ctx* ctx, int64_t n, const double* value);
/* Time-dependent magnitudes for a run of steps: value[r][k] is entry k's magnitude at the time of step r of the
 * next nsm_b200_step call (r = 0 .. n_rows-1; that call must not ask for more steps than rows).  The host
 * evaluates expression(x,y,z,t) exactly where the reference does (src/nimble_boundary_condition_manager.h:166-201)
 * and uploads the table, so that runs of steps stay on the device.  nsm_b200_set_bc_values returns to one row. */
int nsm_b200_set_bc_values_steps(nsm_b200_ctx* ctx, int n_rows, int64_t n, const double* value);
/* Device-evaluated magnitudes (replaces the per-node ExpressionParsing::BoundaryConditionFunctor call of
 * src/nimble_boundary_condition_manager.h:166-201, src/nimble_expression_parser.h:694-760, for expressions
 * whose position-dependent part is IEEE-exact arithmetic).  A program is postfix code over a value stack,
 * code word = op | (arg << 8):
 *   NSM_BCOP_CONST arg -> consts[arg]     NSM_BCOP_X / _Y / _Z -> reference coordinate of the entry's node
 *   NSM_BCOP_SLOT arg  -> slots[step][arg], a scalar the HOST evaluated for that step (the time, and every
 *                         sub-expression of t alone -- cos(t*pi/T) keeps glibc's bits that way)
 *   NSM_BCOP_ENTRYCONST arg -> entry_constants[arg][entry], see nsm_b200_set_bc_entry_constants
 *   ADD SUB MUL DIV FMOD NEG SQRT ABS FLOOR CEIL ROUND, LT LE GT GE EQ AND OR XOR NOT (1.0 / 0.0 results),
 *   SELECT (a b c -> a != 0 ? b : c): all correctly rounded / exact on both sides, hence bit-identical to the
 *   host evaluation.
 * program_of_entry[k] >= 0 makes entry k of the BC table take program's value instead of value[k]; -1 keeps
 * the host-evaluated magnitude.  Per step only n_slots scalars cross the bus, not one value per BC node. */
typedef enum {
  NSM_BCOP_CONST = 0, NSM_BCOP_X = 1, NSM_BCOP_Y = 2, NSM_BCOP_Z = 3, NSM_BCOP_SLOT = 4,
  NSM_BCOP_ADD = 5, NSM_BCOP_SUB = 6, NSM_BCOP_MUL = 7, NSM_BCOP_DIV = 8, NSM_BCOP_FMOD = 9, NSM_BCOP_NEG = 10,
  NSM_BCOP_SQRT = 11, NSM_BCOP_ABS = 12, NSM_BCOP_FLOOR = 13, NSM_BCOP_CEIL = 14, NSM_BCOP_ROUND = 15,
  NSM_BCOP_LT = 16, NSM_BCOP_LE = 17, NSM_BCOP_GT = 18, NSM_BCOP_GE = 19, NSM_BCOP_EQ = 20,
  NSM_BCOP_AND = 21, NSM_BCOP_OR = 22, NSM_BCOP_XOR = 23, NSM_BCOP_NOT = 24, NSM_BCOP_SELECT = 25,
  NSM_BCOP_ENTRYCONST = 26, /* arg -> entry_constants[arg][k]: a value the HOST evaluated once for table entry k */
  NSM_BCOP_COUNT = 27
} nsm_bc_op;
#define NSM_BC_STACK_DEPTH 16
/* Call after nsm_b200_set_bc_table.  program_offsets has n_programs + 1 entries into code; n_entries must equal
 * the BC table length.  Programs are validated here (stack depth, operand indices); n_programs = 0 removes them. */
int nsm_b200_set_bc_programs(nsm_b200_ctx* ctx, int n_programs, const int32_t* program_offsets, const int32_t* code,
                             int n_consts, const double* consts, int n_slots, int64_t n_entries,
                             const int32_t* program_of_entry);
/* Per-entry constants of the programs (NSM_BCOP_ENTRYCONST): values[j][k] = sub-expression j at the node of table
 * entry k.  For sub-trees of the POSITION alone that go through libm or pow (sin(3*x), x^2, exp(-y)): they do not
 * change in time, so the host evaluates them once at set-up with glibc's bits -- where the reference re-evaluates them
 * per node per step (src/nimble_boundary_condition_manager.h:166-201, src/nimble_expression_parser.h:323) -- and the
 * device reads them back every step.  Call after nsm_b200_set_bc_programs; stepping fails with NSM_ERR_ARG while a
 * program names a constant that has not been supplied. */
int nsm_b200_set_bc_entry_constants(nsm_b200_ctx* ctx, int n_constants, int64_t n_entries, const double* values);
/* slots[r][s] for step r of the next nsm_b200_step call (r = 0 .. n_rows-1; that call must not ask for more steps
 * than rows); row 0 also serves nsm_b200_apply_kinematic_bc. */
int nsm_b200_set_bc_slots_steps(nsm_b200_ctx* ctx, int n_rows, int n_slots, const double* slots);
/* Applies the table once at (time_current, time_previous) to the device velocity (row 0 of the magnitudes). */
int nsm_b200_apply_kinematic_bc(nsm_b200_ctx* ctx, double time_current, double time_previous);

/* ---- the explicit step (replaces the loop body of ExplicitTimeIntegrator::Integrate,
 *      src/integrators/explicit_time_integrator.cc:177-278, the contact branch :232-249 included once
 *      nsm_b200_set_contact has been called) -------------------------------------------------------
 * Advances n_steps steps from *time (in/out): per step t_prev = t; t += dt_user; dt = t - t_prev;
 * v += dt/2 a; BC; u += dt v; BC; f_int(u); [f_contact(u);] a = (1/m)(f_int + f_ext [+ f_contact]); v += dt/2 a.
 * BC magnitudes: the row of the last nsm_b200_set_bc_values call; for time-dependent expressions either one
 * host-evaluated row per step (nsm_b200_set_bc_values_steps) or device programs with per-step slots
 * (nsm_b200_set_bc_programs + nsm_b200_set_bc_slots_steps).  store_ipt_last != 0 writes F/sigma on the final step and re-applies the BCs after
 * it, which is what the reference does on an output step (:266-275). */
int nsm_b200_step(nsm_b200_ctx* ctx, int n_steps, double* time, double dt_user, int store_ipt_last);
/* The same loop body on HOST-resident state, i.e. on the reference's Viewify<2> views as ExplicitTimeIntegrator holds
 * them (src/integrators/explicit_time_integrator.cc:131-160): uploads displacement, velocity, acceleration
 * ([n_nodes][3]), advances one step, returns displacement, velocity, acceleration and internal_force in place
 * (internal_force may be NULL: the integrator reads it on output steps only, and a step that does not ask for it moves
 * 6 instead of 7 fields across the bus).
 * The call is PIPELINED over chunks of consecutive node ids (32 by default on meshes of a million nodes or more): chunk
 * c is uploaded and integrated (first half of the step) while chunk c-1's elements run -- every 4-element group whose
 * nodes all lie in the chunks uploaded so far -- and every chunk whose elements have all run is corrected and sent
 * home while later chunks are still travelling up: upload, compute and download overlap (PCIe is full duplex).
 * The dependency ranges come from the connectivity, so the overlap is as good as the mesh numbering is local (lattice
 * or Morton order: a diagonal pipeline; random order: the plain upload -> step -> download schedule).  Per-node and
 * per-element arithmetic and the ORDERED summation order are those of nsm_b200_step, hence the same bits.  Use pinned
 * buffers (nsm_b200_host_alloc).  Contexts with a peer exchange, an internal node renumbering, per-step
 * boundary-condition rows or contact entities (the search needs the whole displacement) take the plain schedule. */
int nsm_b200_step_host(nsm_b200_ctx* ctx, double* time, double dt_user, double* displacement, double* velocity,
                       double* acceleration, double* internal_force);

/* Number of node chunks of the pipelined nsm_b200_step_host: -1 automatic (default), 0 or 1 the plain schedule.  Call
 * before the first nsm_b200_step_host. */
int nsm_b200_set_host_step_chunks(nsm_b200_ctx* ctx, int n_chunks);

/* ---- penalty contact (replaces ContactManager::CreateContactEntities' device arrays, ComputeContactForce and the
 *      contact branch of the explicit loop: src/nimble_contact_manager.cc:184-393, 395-429,
 *      src/contact/serial/arborx_serial_contact_manager.cc:147-196, src/integrators/explicit_time_integrator.cc:232-249) */
/* Contact entities of this context: the skin faces of the primary blocks as quads of node ids (Exodus face order; each
 * becomes four triangles around its centre, src/nimble_contact_manager.cc:1043-1190) with their characteristic
 * lengths, and the contact nodes of the secondary blocks with theirs (the host layer's ContactManager computes both
 * lists as the reference does, :184-330).  From then on every step evaluates the contact force after the internal
 * force and the acceleration is (1/m)(f_int + f_ext + f_contact).  The sum over pairs is atomic (order not fixed, noise
 * ~1e-16) in ATOMIC assembly and ordered like the serial walk in ORDERED assembly.  n_primary_faces == 0 && n_contact_nodes == 0
 * switches contact off again.  Contexts with a peer exchange are refused: for contact across mesh partitions the host
 * layer replicates the contact surface in a second, element-free context (n_blocks == 0, nodes = the surface nodes of
 * all ranks) and calls nsm_b200_contact_force_host on it with the pooled displacements (host/contact_manager.cc,
 * ContactManager::BuildReplicatedSubModel). */
int nsm_b200_set_contact(nsm_b200_ctx* ctx, double penalty_parameter, int64_t n_primary_faces, const int32_t* primary_face_nodes,
                         const double* primary_face_char_len, int64_t n_contact_nodes, const int32_t* contact_node_ids,
                         const double* contact_node_char_len);
/* ContactManager::ComputeContactForce: from the device-resident displacement into the device-resident contact_force
 * (NSM_FIELD_CONTACT_FORCE; zero away from the contact surfaces, ContactManager::GetForces :732-748). */
int nsm_b200_contact_force(nsm_b200_ctx* ctx);
/* The reference call shape (ComputeContactForce(step, debug_output, Viewify<2> contact_force) after the displacement
 * reached the device): uploads displacement [n][3] when it is non-null, evaluates, downloads contact_force [n][3]. */
int nsm_b200_contact_force_host(nsm_b200_ctx* ctx, const double* displacement, double* contact_force);
/* Counters of the last evaluation: stats[0] node-face pairs enforced, [1] pairs that passed the bounding-box test,
 * [2] triangles in contact (numActiveContactFaces, :692-702), [3] nodes in contact (numActiveContactNodes), [4] ORDERED
 * assembly only: pairs that did not fit the ordered lists (room for four pairs per contact node) and were added
 * atomically -- 0 means the contact force was summed in the serial order (contact nodes ascending, triangles ascending,
 * facet nodes before the node) and is bit-reproducible. */
int nsm_b200_contact_stats(nsm_b200_ctx* ctx, int64_t stats[5]);
/* ContactEntity::contact_status() of every entity after the last evaluation (what ContactVisualizationWriteStep copies
 * back with its entities, src/nimble_contact_manager.cc:596-602, 640-668): face_status [4 * n_faces], triangle k of
 * face f at 4 f + k (the entity order of CreateContactNodesAndFaces, :1043-1190), node_status [n_contact_nodes] in
 * the order of set_contact; 1 = in contact.  Either pointer may be NULL.  All zero before the first evaluation. */
int nsm_b200_contact_status(nsm_b200_ctx* ctx, unsigned char* face_status, unsigned char* node_status);

/* ---- element data / derived output (replaces ModelData::GetElementDataNew + Block::ComputeDerivedElementData,
 *      src/nimble_block.cc:438-497; HexElement::ComputeVolumeAverage, src/nimble_element.h:343-392) --- */
int nsm_b200_get_element_data(nsm_b200_ctx* ctx, int block_id, double* out /*[n_elem][8][stride]*/);
/* doubles per integration point of a block's records: 15 + the state variables of its material (label order of
 * Block::GetDataLabelsAndLengths, src/nimble_block.cc:84-108: F 9, sigma 6, then the state scalars) */
int nsm_b200_element_data_stride(const nsm_b200_ctx* ctx, int block_id);
/* ModelData::UpdateStates (src/nimble_model_data.h:104-107; end of the loop body, explicit_time_integrator.cc:277):
 * the records written by the last force evaluation become the N records of the next one.  The swap is applied when
 * that next evaluation starts, so nsm_b200_get_element_data keeps returning the most recently computed records
 * (what the reference writes on an output step, before its swap).  nsm_b200_step does this after every step;
 * callers sequencing nsm_b200_internal_force themselves call it where the reference calls UpdateStates. */
int nsm_b200_update_states(nsm_b200_ctx* ctx);
/* Overwrites a block's records from a host array [n_elem][8][stride]: previous = 0 the current (N+1) records,
 * previous != 0 the N records the next force evaluation reads (blocks with state variables only).  This is how
 * Block::ComputeInternalForce receives the caller's elem_data_n (src/nimble_block.cc:297-337), and how a run is
 * resumed from saved element data. */
int nsm_b200_set_element_data(nsm_b200_ctx* ctx, int block_id, int previous, const double* in);
/* The N records of a block whose material has state variables (element_data_n of the reference). */
int nsm_b200_get_element_data_previous(nsm_b200_ctx* ctx, int block_id, double* out /*[n_elem][8][stride]*/);
/* out [1 + stride][n_elem]: volume, then volume averages of F (9), sigma (6) and the state scalars in storage order. */
int nsm_b200_derived_element_data(nsm_b200_ctx* ctx, int block_id, double* out);
/* Selected integration-point components of one block, split on the device: out[k][e] = ipt[e][offsets[k]],
 * offsets in 0..8*stride-1 = stride * point + field (the reference's per-element label order, src/nimble_block.cc:84-108).
 * Replaces the host loop of ModelData::WriteExodusOutput over GetElementDataNew (src/nimble_model_data.cc:557-596):
 * only the requested columns cross the bus (8 B per element and component instead of 960 B per element). */
int nsm_b200_get_element_components(nsm_b200_ctx* ctx, int block_id, int n_components, const int32_t* offsets, double* out);

/* Full integration-point records of a LIST of elements of one block (file-order element indices), gathered on
 * the device: out[i] = the [8][15] record of elements[i].  For checks and probes on meshes whose whole
 * [n_elem][8][15] array (61 GB at 64 M elements) should not cross the bus; same data as
 * ModelData::GetElementDataNew (src/nimble_model_data.cc:557-596) restricted to those elements. */
int nsm_b200_get_element_data_subset(nsm_b200_ctx* ctx, int block_id, int64_t n, const int64_t* elements, double* out);

/* ---- shared-node exchange over NVLink peer memory (replaces VectorCommunicator::VectorReduction /
 *      ReductionClique_t MPI_Iallreduce, src/nimble_vector_communicator.h:104-157,
 *      src/nimble.mpi.rank_clique_reducer.h:130-257) -----------------------------------------------
 * One context per rank (process-per-GPU or thread-per-GPU).  The host layer discovers the nodes this rank
 * shares with each other rank from the global node ids, exactly as GenerateReductionInfo does
 * (src/nimble.mpi.reduction.cc:50-123), and passes, for every peer rank, the LOCAL ids of the nodes shared
 * with that peer sorted by GLOBAL id (both sides sort alike, :114-120).  Receive buffers are exported as
 * opaque blobs that the host layer passes between ranks by any transport (torch.distributed, MPI, files,
 * or a plain memcpy between threads).  Sums are formed in ascending rank order on every holder, so all
 * replicas of a shared node carry bit-identical values. */
#define NSM_COMM_HANDLE_BYTES 192
int nsm_b200_comm_init(nsm_b200_ctx* ctx, int rank, int world_size, int n_peers, const int32_t* peer_ranks,
                       const int64_t* pair_offsets /*[n_peers+1]*/, const int32_t* pair_local_nodes);
int nsm_b200_comm_export(nsm_b200_ctx* ctx, unsigned char handle[NSM_COMM_HANDLE_BYTES]);
int nsm_b200_comm_attach(nsm_b200_ctx* ctx, int peer_rank, const unsigned char handle[NSM_COMM_HANDLE_BYTES]);
/* All peers attached: build the device tables; from here on lumped mass and internal force include the
 * shared-node sum.  Every rank must be past comm_attach of all its peers before any rank steps. */
int nsm_b200_comm_ready(nsm_b200_ctx* ctx);

/* Ranks that SHARE one GPU (several contexts on a device: the multi-rank path on a single-GPU box) cannot wait for each
 * other inside a kernel.  With a host barrier set, every exchange becomes pack -> host waits for its own pack ->
 * barrier(arg) (must return once ALL ranks of the run have called it for this exchange) -> rank-ordered unpack; the
 * sums are the same bits.  One GPU per rank keeps the in-kernel wait (no host round trip per step).  Call on every
 * rank before the first exchange; barrier = NULL restores the in-kernel wait. */
int nsm_b200_comm_set_host_barrier(nsm_b200_ctx* ctx, void (*barrier)(void*), void* arg);
/* How long a rank waits inside the step for a peer's shared-node data before the step fails with NSM_ERR_COMM
 * (default 20 s; the reference's MPI_Wait has no bound, src/nimble.mpi.rank_clique_reducer.h:230-257).  Raise it
 * when ranks may drift apart by more than that between two steps (a rank blocked in file output). */
int nsm_b200_comm_set_timeout(nsm_b200_ctx* ctx, double seconds);

/* ---- measurement helpers (CUDA events on the context stream) ---------------------------------- */
int nsm_b200_timer_start(nsm_b200_ctx* ctx);
int nsm_b200_timer_stop(nsm_b200_ctx* ctx, float* milliseconds);
/* Kernels launched by this context since creation (bench.py reports the delta as gpu_launches). */
int64_t nsm_b200_launch_count(const nsm_b200_ctx* ctx);
/* Average device time of the element (internal-force) kernel launches since the last reset, measured
 * with CUDA events around each launch when profiling is switched on. */
int nsm_b200_profile(nsm_b200_ctx* ctx, int enable);
int nsm_b200_profile_read(nsm_b200_ctx* ctx, double* elem_kernel_ms_avg, double* node_kernel_ms_avg, int64_t* n_launches);
/* the contact evaluation's share of the same profiled steps (the reference times it as its own region "Contact",
 * src/integrators/explicit_time_integrator.cc:233-236); node_kernel_ms_avg above does not include it */
int nsm_b200_profile_read_contact(nsm_b200_ctx* ctx, double* contact_ms_avg);
/* Integration points that left the branch-free arithmetic window since creation and were recomputed with the
 * plain IEEE operators (same bits, slower; see csrc/hex8_math.cuh).  Wraps at 2^32. */
int64_t nsm_b200_cold_points(nsm_b200_ctx* ctx);
/* FP64 pipe micro-benchmark: sustained DADD+DMUL (no FMA) and DFMA issue rates in 1e12 lane-ops/s. */
int nsm_b200_fp64_peak(nsm_b200_ctx* ctx, double* dadd_dmul_tops, double* dfma_tops);
/* The same DADD+DMUL stream back to back for `seconds` (<= 30): the rate of the last quarter, i.e. what an FP64-bound
 * run of that length can sustain once power capping has settled the SM clock (the burst figure above is the
 * denominator for a kernel timed alone, this one for a kernel timed inside a long run). */
int nsm_b200_fp64_peak_sustained(nsm_b200_ctx* ctx, double seconds, double* dadd_dmul_tops);

/* Build-time description of the element kernels of THIS binary (JSON text): for every element_force_kernel
 * instance the static instruction mix of one warp pass over 4 elements, read off the SASS of the object the
 * library was linked from (scripts/sass_hot_loop.py): "dp" (DADD + DMUL + DFMA + DSETP), "other",
 * "dp_lane_instr_per_element" = dp * 8, registers, stack bytes, and "source_sha", a hash of the kernel sources.
 * bench.py derives its FP64-pipe figures from this text, so they cannot go stale against the code. */
const char* nsm_b200_kernel_info(void);

#ifdef __cplusplus
}
#endif
#endif /* NSM_B200_H */
