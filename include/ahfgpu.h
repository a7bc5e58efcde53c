/*
 * ahfgpu.h -- C-ABI of libahfgpu.so: the B200 (sm_100a) implementation of AHF's particle hot path.
 *
 * Plain C, plain pointers and sizes, int status returns (0 = ok, <0 = error, text via ahfgpu_last_error()).
 * The reference (NegriAndrea/AHF, C99) has no plugin/FFI layer; these entry points replace ordinary C calls
 * inside its own translation units (INTEGRATION.md shows the call-site patch).  Each entry point names the
 * reference interface it replaces (file:line relative to the reference tree).
 *
 * Ownership: every pointer passed in is HOST memory owned by the caller; outputs are written into caller
 * buffers (query sizes first) -- nothing returned here has to be freed by the caller, and nothing the
 * reference later free()s is allocated here.  All calls are made from one host thread (the reference's main
 * thread); the library uses its own CUDA streams internally.
 * Units: the reference's internal units (positions in box units [0,1), momenta a^2 dx/dt / (box/t_unit),
 * masses in units of simu.pmass).
 */
#ifndef AHFGPU_H
#define AHFGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AHFGPU_NSCAL    64   /* scalar slots per halo, layout below */
#define AHFGPU_NPROFCOL 25   /* profile columns per radial bin     */

typedef struct ahfgpu_ctx ahfgpu_ctx;

/* POD copy of the run parameters the path needs: simu.* (src/tdef.h:315-411, filled at src/startrun.c:465-713)
 * and the unit factors / cosmology scalars of src/libahf/ahf_halos.c:199-221. */
typedef struct {
  int32_t device;          /* CUDA device ordinal                                                        */
  int32_t lgrid_dom;       /* simu.NGRID_DOM  (AHF.input LgridDomain, power of two <= 2^21)               */
  int32_t lgrid_max;       /* simu.NGRID_MAX  (AHF.input LgridMax; clipped to 2^21)                       */
  int32_t min_part;        /* simu.AHF_MINPART (NminPerHalo)                                              */
  double  nth_dom;         /* simu.Nth_dom    (NperDomCell)                                               */
  double  nth_ref;         /* simu.Nth_ref    (NperRefCell)                                               */
  double  vesc_tune;       /* simu.AHF_VTUNE  (VescTune)                                                  */
  double  r_fac, x_fac, v_fac, m_fac, rho_fac, phi_fac;   /* ahf_halos.c:199-205                           */
  double  hubble;          /* calc_Hubble(a)  ahf_halos.c:213                                             */
  double  ovlim;           /* global.ovlim    ahf_halos.c:212,219                                         */
  double  rho_vir;         /* global.rho_vir  ahf_halos.c:216,221                                         */
} ahfgpu_params;

const char *ahfgpu_last_error(void);
int  ahfgpu_device_count(void);

/* after startrun() (src/main.c:128) / before the final frees (src/main.c:671-707).  ahfgpu_warmup(device) may be called earlier, from a
 * helper thread at program start: it creates the CUDA context and loads the kernels while the host parses its parameter file */
int  ahfgpu_warmup(int32_t device);
int  ahfgpu_init(ahfgpu_ctx **ctx, const ahfgpu_params *par);
int  ahfgpu_set_params(ahfgpu_ctx *ctx, const ahfgpu_params *par);
int  ahfgpu_finalize(ahfgpu_ctx *ctx);

/* ---- K1 + K2 ----------------------------------------------------------------------------------------------
 * Replaces the key loop + qsort of src/main.c:343-356 (sfc_curve_calcKey, src/libsfc/sfc_curve.c:87;
 * cmp_sfckey_part, src/libutility/specific.c:1116).
 *
 * ahfgpu_sfc_sort_particles works on the reference's own AoS: `part` is `n` records of `stride` bytes
 * (sizeof(struct particle), src/tdef.h:36-77; a multiple of 8); off_* are byte offsets of pos[3] (float),
 * mom[3] (float), sfckey (uint64), id (uint64) and -- or -1 when the build has no such field -- weight
 * (float) and u (float).  On return the HOST array is sorted ascending by key with sfckey filled (the first
 * member `ll` is zeroed), exactly what ahf_halos.c:3361 / ahf_io.c:1136 index later, and the sorted particles
 * stay resident on the device for the calls below.  Tie order between equal keys is by input position.
 *
 * ahfgpu_sfc_sort_soa is the same for separate arrays; keys_out (n) / order_out (n, input position of the
 * particle now at sorted offset i) may be NULL.                                                            */
int  ahfgpu_sfc_sort_particles(ahfgpu_ctx *ctx, void *part, uint64_t n, uint32_t stride, int32_t off_pos,
                               int32_t off_mom, int32_t off_key, int32_t off_id, int32_t off_weight, int32_t off_u);
int  ahfgpu_sfc_sort_soa(ahfgpu_ctx *ctx, const float *pos3, const float *mom3, const float *weight, const float *u,
                         uint64_t n, uint64_t *keys_out, uint32_t *order_out);
/* ahfgpu_sfc_sort_soa_async: same result as ahfgpu_sfc_sort_soa, but the host->device copies run on a second stream and
 * overlap the work: keys are computed chunk by chunk as the positions land, and the momenta (and u) travel while the sort
 * and ahfgpu_build_amr run -- only ahfgpu_construct_halos needs them and waits for them on the device.  The call returns
 * when the sorted positions and keys are resident.  CONTRACT: mom3 and u must stay valid and unchanged until the next
 * ahfgpu_construct_halos, ahfgpu_synchronize or ahfgpu_finalize on this context has returned (pos3 and weight are free on
 * return); pinned host memory is needed for the copies to overlap at all.                                             */
int  ahfgpu_sfc_sort_soa_async(ahfgpu_ctx *ctx, const float *pos3, const float *mom3, const float *weight, const float *u, uint64_t n);
/* Same as ahfgpu_sfc_sort_soa but split in two so that the sort can be timed with its input already in HBM:
 * ahfgpu_upload_soa copies the unsorted arrays to the device, ahfgpu_sfc_sort_resident runs keys + sort + gather on
 * them (repeatable: the unsorted copy is kept).                                                                */
int  ahfgpu_upload_soa(ahfgpu_ctx *ctx, const float *pos3, const float *mom3, const float *weight, const float *u, uint64_t n);
int  ahfgpu_sfc_sort_resident(ahfgpu_ctx *ctx);
/* keys only (no sort, nothing stays resident): sfc_curve_calcKey(SFC_CURVE_HILBERT, x, y, z, bits) per particle */
int  ahfgpu_hilbert_keys(ahfgpu_ctx *ctx, const float *pos3, uint64_t n, uint32_t bits, uint64_t *keys_out);

/* ---- D / F / R / L ------------------------------------------------------------------------------------------
 * Replaces gen_domgrids + ll + zero_dens + assign_npart + gen_AMRhierarchy (src/main.c:616-648;
 * src/libamr_serial/generate_grids.c:24-432, lltools.c:32, density.c:238-492, refine_grid.c:903, relink.c:31)
 * on the resident sorted particles.  The hierarchy stays on the device; the query calls copy it out.        */
int  ahfgpu_build_amr(ahfgpu_ctx *ctx);
int  ahfgpu_amr_nlevels(ahfgpu_ctx *ctx);
/* iout[0]=l1dim [1]=ncell [2]=particles deposited on the level [3]=particles finally owned by it;
 * dout[0]=critdens [1]=masstopartdens  (gridls fields, src/tdef.h:189-235) */
int  ahfgpu_amr_level_header(ahfgpu_ctx *ctx, int32_t lev, int64_t *iout, double *dout);
/* Cells in the reference's traversal order (z, y, x ascending).  Any pointer may be NULL.
 *   dens     node.dens (float, number density contrast, src/libamr_serial/density.c:398,480)
 *   runflags bit0/1 first/last node of its nquad, bit2/3 first/last row of its cquad, bit4/5 first/last plane
 *            of its pquad -- enough to rebuild the pquad/cquad/nquad runs of src/tdef.h:143-183
 *   interior test_tsc() of src/libamr_serial/get_nnodes.c:41-51
 *   mark     0 untouched / 1 refined / 2 ghost pair (src/libamr_serial/refine_grid.c:231-250)
 *   count    particles linked to the node when the level was deposited                                      */
int  ahfgpu_amr_level_get(ahfgpu_ctx *ctx, int32_t lev, int32_t *x, int32_t *y, int32_t *z, float *dens,
                          uint8_t *runflags, uint8_t *interior, uint8_t *mark, int32_t *count);
/* NEXT-1 of SURVEY 8f (first step past the hot path): the patch colouring of ahf_gridinfo (src/libahf/ahf_gridinfo.c:236-577) on the
 * device.  iso[ncell]: index of the isolated refinement (6-connected patch over the neighbours the reference's search sees) every
 * cell of level `lev` belongs to, numbered as the reference numbers them (spatialRefIndex[colour].isoRefIndex, :751-775: in the
 * order of their first cell); *niso: their number (numIsoRef[lev - min_ref]); periodic3[3*i + d] (capacity 3*ncell bytes, may be
 * NULL): SRINDEX.periodic.x/y/z of patch i (testBound, :1090-1118).  Cell order as in ahfgpu_amr_level_get. */
int  ahfgpu_amr_patches(ahfgpu_ctx *ctx, int32_t lev, int32_t *iso, int64_t *niso, uint8_t *periodic3);
/* NEXT-2 of SURVEY 8f, first half: RefCentre (src/libahf/ahf_halos.c:935-1620) as segmented reductions over the patch labels of
 * level `lev`.  stats[i*18 + k] for isolated refinement i (numbered as in ahfgpu_amr_patches): 0 numNodes, 1 numParts (particles the
 * level finally owns), 2-4 centre (= centre of mass of those particles: the shipped define.h:101 sets AHFcomcentre; geometric centre
 * where there are none), 5 maxDens, 6-8 centreGEOM, 9-11 centreDens (density weighted, with the fall-backs of :1248-1350),
 * 12-17 x.min x.max y.min y.max z.min z.max (periodic refinements are cut at boundRefDiv, so max < min there, :1400-1612).
 * *niso: number of refinements; stats may be NULL (count only), stats_cap = its capacity in refinements. */
int  ahfgpu_amr_patch_stats(ahfgpu_ctx *ctx, int32_t lev, int64_t *niso, double *stats, int64_t stats_cap);
/* NEXT-2, second half -- HOST code (no device work): the refinement tree and the halo seeds from the tables of
 * ahfgpu_amr_patch_stats.  Replaces analyseRef (src/libahf/ahf_halos.c:1652-2300) and spatialRef2halos (:2405-3058), default switches
 * of the shipped define.h (PARDAU_PARTS, AHFcomcentre).  In: nlev coloured levels (level 0 = ahf.min_ref), niso[nlev] refinements per
 * level, stats = their 18-column rows concatenated level by level, max_gather_rad = MaxGatherRad / boxsize.  Out, per refinement row
 * (any pointer may be NULL): daughter = main-branch refinement on the next level or -1 (SPATIALREF.daughter), close_ref_dist
 * (.closeRefDist), substructure lists as CSR (sub_offset[rows+1], sub[] = indices on the next level, in the reference's order).
 * Out, per halo in the order of the reference's halos[] array: centre (HALO.pos), gatherRad, npart, hostHalo -- the inputs of
 * ahfgpu_construct_halos.  *nhalo is always set; buffers are checked against sub_cap / halo_cap. */
int  ahfgpu_tree_halos(int32_t nlev, const int64_t *niso, const double *stats, double max_gather_rad,
                       int32_t *daughter, double *close_ref_dist, int64_t *sub_offset, int32_t *sub, int64_t sub_cap,
                       int64_t *nhalo, double *halo_pos3, double *halo_gather_rad, int64_t *halo_npart, int32_t *halo_host,
                       int64_t halo_cap);
/* the same with what the sub-halo re-hash needs (ahfgpu_catalogue_write): per halo hostHaloLevel (coloured level that opened the
 * sub-halo, ahf_halos.c:2699 / :2863; -1 for a field halo) and the substructure lists of the halos[] array as spatialRef2halos leaves
 * them (CSR over the haloes, halo_sub_offset[nhalo + 1]; entries in the reference's order).  Capacity of halo_sub: halo_sub_cap. */
int  ahfgpu_tree_halos_ex(int32_t nlev, const int64_t *niso, const double *stats, double max_gather_rad,
                          int32_t *daughter, double *close_ref_dist, int64_t *sub_offset, int32_t *sub, int64_t sub_cap,
                          int64_t *nhalo, double *halo_pos3, double *halo_gather_rad, int64_t *halo_npart, int32_t *halo_host,
                          int64_t halo_cap, int32_t *halo_host_level, int64_t *halo_sub_offset, int32_t *halo_sub, int64_t halo_sub_cap);
/* per particle (sorted offset): deepest level that owns it (node.ll membership after all relinks) and its cell
 * index on every level it reached: cell_of[lev*n + i] = index into the level's cell list or -1 */
int  ahfgpu_amr_particle_levels(ahfgpu_ctx *ctx, int8_t *owner_level, int32_t *cell_of, int32_t nlev_cap);

/* ---- G / U / P ----------------------------------------------------------------------------------------------
 * Replaces the OpenMP loop over ahf_halos_sfc_constructHalo (src/libahf/ahf_halos.c:504-510;
 * ahf_halos_sfc.c:118-169): gather, radial sort, virial cut, unbinding, virial cut, profiles.
 * In: per halo centre[3] (HALO.pos), gather_rad (HALO.gatherRad), seed_npart (HALO.npart; 0 = skip,
 * ahf_halos_sfc.c:122).  Results stay on the device until fetched:
 *   scal     nhalo x AHFGPU_NSCAL doubles:
 *              5 n_gathered, 6 n after 1st virial cut, 7 n after unbinding, 8 n after 2nd virial cut, 9 npart,
 *              10 M_vir, 11 R_vir, 12 ovdens, 13 Phi0, 14-16 vel, 17 sigV, 18 v_esc2, 19 V2_max, 20 R_max, 21 r2,
 *              22 lambda, 23 lambdaE, 24 Ekin, 25 Epot, 26 SurfP, 27-29 pos_com, 30 com_offset, 31-33 pos_mbp,
 *              34-36 vel_mbp, 37 mbp_offset, 38-40 AngMom, 41-43 axis, 44-52 E1,E2,E3, 53 fMhires, 54 cNFW,
 *              55 cR1, 56 R1, 57 nbins            (HALO fields, src/tdef.h:692-789)
 *   members  radius-sorted offsets into the sorted particle array (HALO.ipart), CSR via member_offset
 *   prof     per halo nbins x AHFGPU_NPROFCOL (column-major, col*nbins+bin): npart, r, nvpart, ovdens, dens,
 *            v2_circ, v_esc2, sig_v, Ekin, Epot, Lx, Ly, Lz, axis1, E1x, E1y, E1z, axis2, E2x, E2y, E2z, axis3,
 *            E3x, E3y, E3z   (HALOPROFILE, src/tdef.h:591-688), CSR via prof_offset (in bins)                 */
int  ahfgpu_construct_halos(ahfgpu_ctx *ctx, int64_t nhalo, const double *centre3, const double *gather_rad,
                            const int64_t *seed_npart);
int  ahfgpu_halo_sizes(ahfgpu_ctx *ctx, int64_t *total_members, int64_t *total_bins);
int  ahfgpu_halo_fetch(ahfgpu_ctx *ctx, double *scal, int64_t *member_offset, int64_t *members,
                       int64_t *prof_offset, double *prof);
/* Optional: a PINNED host buffer for the member lists, registered before ahfgpu_construct_halos.  The lists are final once the
 * unbinding is done (src/libahf/ahf_halos.c:3361ff reads them only after the halo loop), so the library sends them home on its copy
 * stream while the profiles are still being computed; ahfgpu_halo_fetch called with the same `members` pointer then only waits for that
 * copy.  capacity in entries; a call whose lists do not fit falls back to the copy inside ahfgpu_halo_fetch.  NULL unregisters. */
int  ahfgpu_halo_members_buffer(ahfgpu_ctx *ctx, int64_t *members_pinned, int64_t capacity);
/* -DGAS_PARTICLES build of the reference (particles carry `u`): the per-species blocks of HaloProfiles
 * (src/libahf/ahf_halos.c:4420-4582, :4712-4715, :5020-5181).
 *   species       nhalo x 64 doubles: HALO.gas_only at 0, HALO.stars_only at 32 (SPECIESPROP, src/tdef.h:560-587):
 *                 0 npart, 1 Mass, 2-4 pos_com, 5-7 pos_mbp, 8-10 vel, 11 lambda, 12 lambdaE, 13-15 AngMom, 16-18 axis,
 *                 19-27 E1 E2 E3, 28 Ekin, 29 Epot
 *   prof_species  per halo nbins x 3 (column-major, CSR via prof_offset): HALOPROFILE.M_gas, .M_star, .u_gas        */
int  ahfgpu_halo_fetch_species(ahfgpu_ctx *ctx, double *species, double *prof_species);

/* ---- several GPUs working on ONE box -------------------------------------------------------------------------------
 * Replaces the reference's MPI mode: loadbalance_update / local_equalpart (src/libutility/loadbalance.c:206,383: histogram of the
 * particles over the Hilbert cells of LevelDomainDecomp, MPI_Allreduce :480, equal-particle key ranges), comm_dist_part
 * (src/comm.c:104-316: every particle to the owner of its key range) and comm_dist_part_ahf (src/comm.c:324ff with
 * sfc_boundary_2_get, src/libsfc/sfc_boundary.c:100-144: duplication of the boundary-shell particles on the neighbours).
 * One context per GPU; every context is given a communicator and the particles ITS process read (any subset, any order):
 *   ahfgpu_comm_nccl_unique_id / ahfgpu_comm_init_nccl : one process per GPU, NCCL over NVLink (id: 128 bytes made on rank 0 and handed
 *                                to all ranks by the caller -- MPI_Bcast, torch.distributed, a file ...)
 *   ahfgpu_comm_local_group_create / ahfgpu_comm_init_local : several contexts of ONE process, each driven by its own host thread (they
 *                                may share a device); for tests of the decomposition on a box with one GPU
 *   ahfgpu_upload_soa + ahfgpu_slab_distribute : keys, block histogram (all-reduce), equal-particle Hilbert ranges, ONE personalised
 *                                exchange that sends every particle to its owner and to the ranks whose range lies within `ghost_width`
 *                                (box units; at least 8 domain cells, the reach of the hierarchy construction, and >= the largest
 *                                gathering radius) of it, ONE sort.  id_base = global input index of the rank's first particle (ranks
 *                                in rank order; all indices < 2^32).  decomp_bits = LevelDomainDecomp (0: blocks of 4 domain cells).
 * Afterwards ahfgpu_build_amr builds every level over the rank's own cells + ghost shell (results are exact on the own cells --
 * ahfgpu_amr_level_owned -- and identical to a single-GPU run, bit for bit), and ahfgpu_construct_halos serves the haloes whose
 * centre the rank owns (ahfgpu_slab_owner_of); member lists then hold GLOBAL INPUT INDICES instead of sorted offsets.
 *   ahfgpu_slab_info : iout[12] = rank, nranks, resident particles (own + ghost), first own, one past last own (sorted offsets),
 *                      particles of the box, decomp bits, shell thickness in blocks, first own block, one past last own block,
 *                      levels of the whole box, collectives since distribute; dout[4] = ghost width, ms inside device collectives
 *                      (CUDA events), bytes received by them, 0
 *   ahfgpu_particle_ids : per resident sorted particle its input position (single GPU) / global input index (slab)            */
int  ahfgpu_comm_nccl_unique_id(void *id128);
int  ahfgpu_comm_init_nccl(ahfgpu_ctx *ctx, int32_t rank, int32_t nranks, const void *id128);
void *ahfgpu_comm_local_group_create(int32_t nranks);
int  ahfgpu_comm_local_group_destroy(void *group);
int  ahfgpu_comm_local_group_abort(void *group);      /* wakes every rank waiting in a collective with an error (a peer failed) */
int  ahfgpu_comm_init_local(ahfgpu_ctx *ctx, int32_t rank, void *group);
int  ahfgpu_slab_distribute(ahfgpu_ctx *ctx, uint64_t id_base, double ghost_width, int32_t decomp_bits);
int  ahfgpu_slab_info(ahfgpu_ctx *ctx, int64_t *iout, double *dout);
int  ahfgpu_slab_owner_of(ahfgpu_ctx *ctx, int64_t n, const double *pos3, int32_t *owner);
int  ahfgpu_amr_level_owned(ahfgpu_ctx *ctx, int32_t lev, uint8_t *owned);
int  ahfgpu_particle_ids(ahfgpu_ctx *ctx, uint32_t *ids);
/* the same after ahfgpu_sfc_sort_soa_async, on a stream of its own (device -> host while the momenta come in and the hierarchy is built);
 * `ids` (pinned) is complete when ahfgpu_construct_halos has returned */
int  ahfgpu_particle_ids_async(ahfgpu_ctx *ctx, uint32_t *ids);
/* lower-level pieces kept for callers that place particles themselves:
 *   ahfgpu_set_global_count : N of the whole box (masstopartdens = L^3 / N, generate_grids.c:66-69)
 *   ahfgpu_adopt_sorted     : make caller-owned DEVICE arrays (float4 pos+weight, float4 mom+u, u64 keys; key sorted) the
 *                             resident particle set (not freed by the library)
 *   ahfgpu_device_ptr       : device pointers of the resident sorted set ("pos4","mom4","keys")
 *   ahfgpu_sfc_sort_device4 : keys + sort + gather of particles that already sit on the device as float4 (x,y,z,weight) /
 *                             float4 (px,py,pz,u) arrays                                                                  */
int  ahfgpu_sfc_sort_device4(ahfgpu_ctx *ctx, const void *pos4_dev, const void *mom4_dev, uint64_t n, int32_t has_weight, int32_t has_u);
int  ahfgpu_set_global_count(ahfgpu_ctx *ctx, uint64_t n_total);
int  ahfgpu_adopt_sorted(ahfgpu_ctx *ctx, const void *pos4_dev, const void *mom4_dev, const void *keys_dev, uint64_t n,
                         int32_t has_weight, int32_t has_u);
void *ahfgpu_device_ptr(ahfgpu_ctx *ctx, const char *name);

/* ---- NEXT-4 of SURVEY 8f: bulk snapshot ingest --------------------------------------------------------------------------------------
 * Replaces, for single-file GADGET-1/2 snapshots whose particle masses are in the header (no MASS block) and that hold no gas (no U
 * block) -- BASELINE.json configs 1-4 --, io_gadget_readpart_raw (src/libio/io_gadget.c:427-568: one fread per value) and
 * io_gadget_scale_particles (:857-995): the POS / VEL / ID blocks are read with one pread each into pinned memory and the extreme
 * positions, the shift, the box check and the conversion to internal units run on the device in the reference's float32 arithmetic,
 * so positions, momenta and keys are bit-identical to the reference's after startrun().  Afterwards the context holds the unsorted
 * particles like after ahfgpu_upload_soa: continue with ahfgpu_sfc_sort_resident.  posscale / weightscale = GADGET_LUNIT / GADGET_MUNIT
 * of AHF.input (<= 0: 1).  ids_out (n, may be NULL; query n with a first call): the ID block.  info[24]: 0 particles, 1 boxsize
 * (simu.boxsize), 2 expansion, 3 omega0, 4 lambda0, 5 pmass, 6-8 shift applied, 9 scale_pos, 10 scale_mom, 11 GADGET version,
 * 12 byte swapped, 13 hubble parameter, 14 ms reading + staging (host wall clock), 15 ms on the device (upload + kernels, CUDA events),
 * 16-18 / 19-21 smallest / largest raw position per axis (io_gadget_t minpos / maxpos, io_gadget.c:1320-1347), 22 header boxsize after
 * the extent check (:879-896, file units), 23 reserved.  Unsupported files are refused with an error, never read partially.
 * ahfgpu_ingest_prefetch(path): the host-only half (header checks + the three block reads into pageable memory), kept until the next
 * ahfgpu_ingest_gadget of the same path; needs no CUDA context, so a host program can read while its context is still being created.
 * ahfgpu_input_peek: one particle of the unsorted input set (the reference logs the first and last one, startrun.c:378-431).
 * ahfgpu_particles_get: the resident SORTED particles (float4 x,y,z,weight / float4 px,py,pz,u) copied to the host (tests).       */
int  ahfgpu_ingest_gadget(ahfgpu_ctx *ctx, const char *path, double posscale, double weightscale, uint64_t *ids_out, double *info);
int  ahfgpu_particles_get(ahfgpu_ctx *ctx, float *pos4, float *mom4);
int  ahfgpu_ingest_prefetch(const char *path);
int  ahfgpu_input_peek(ahfgpu_ctx *ctx, uint64_t index, float *pos3, float *mom3);

/* ---- NEXT-3 of SURVEY 8f: what ahf_halos() does after the halo loop -- HOST code, no device work --------------------------------------
 * Replaces the sub-halo re-hash with the final radii (src/libahf/ahf_halos.c:550-640, check_subhalo :5900-5916), the ordering by particle
 * number (:720-741, Numerical Recipes' indexx, libutility/general.c:1093: the order of haloes with equal counts is reproduced) and the
 * writers of the default build (libahf/ahf_io.c: <fprefix>.AHF_halos :2209-2345, .AHF_profiles :637-826, .AHF_substructure :429-503,
 * .AHF_particles :1094-1175), byte for byte.  All arrays are in the order of the reference's halos[] array (ahfgpu_tree_halos):
 *   scal / member_off / members / prof_off / prof / species / prof_species  as delivered by ahfgpu_halo_fetch(_species); members index
 *                  part_id / part_u / part_weight (any numbering the caller likes, e.g. the input index of ahfgpu_particle_ids);
 *   pos3           HALO.pos (the centres handed to ahfgpu_construct_halos);
 *   host, host_level, sub_off / sub   hostHalo, hostHaloLevel and subStruct[] as spatialRef2halos leaves them (ahfgpu_tree_halos_ex);
 *   part_u         GAS_PARTICLES build: type column = u >= 0 ? 0 : -u;  part_weight: MULTIMASS build without gas: (int)weight;  both NULL: 1;
 *   flags bit 0    multi-species columns (the -DMULTIMASS -DGAS_PARTICLES build: gas / star blocks, M_gas M_star U_gas profile columns).
 * fprefix = "<outfile_prefix>.z<redshift %.3f>" (ahf_halos.c:246-256); NULL: no files, only the outputs below.
 * Optional outputs (nhalo each): host_out = hostHalo after the re-hash, nsub_out = numSubStruct after it, rank_out = the halo's ID
 * (its position in the written order, counted over ALL haloes like the reference's j).                                            */
typedef struct ahfgpu_catalogue_in {
  int64_t         nhalo;
  const double   *scal;
  const double   *pos3;
  const int64_t  *member_off;
  const int64_t  *members;
  const int64_t  *prof_off;
  const double   *prof;
  const double   *species;
  const double   *prof_species;
  const int32_t  *host;
  const int32_t  *host_level;
  const int64_t  *sub_off;
  const int32_t  *sub;
  const uint64_t *part_id;
  const float    *part_weight;
  const float    *part_u;
  double          x_fac, r_fac, v_fac, m_fac, rho_fac, phi_fac, u_fac, rho_vir, pmass;
  int32_t         min_part;
  int32_t         flags;
} ahfgpu_catalogue_in;
int  ahfgpu_catalogue_write(const char *fprefix, const ahfgpu_catalogue_in *in, int32_t *host_out, int32_t *nsub_out, int64_t *rank_out);

/* ---- measurement hooks (bench.py): milliseconds of the last call, by stage, measured with CUDA events on
 * the library's stream.  names: "h2d","keys","sort","gather","d2h","deposit","flag","refine","relink",
 * "halo_gather","halo_sort","halo_unbind","halo_profiles", ... ; returns <0 for an unknown name.             */
double  ahfgpu_stage_ms(ahfgpu_ctx *ctx, const char *name);
/* CUDA events on the library's own stream (slot 0..15): record, then elapsed milliseconds between two slots */
int     ahfgpu_event_record(ahfgpu_ctx *ctx, int32_t slot);
double  ahfgpu_event_elapsed_ms(ahfgpu_ctx *ctx, int32_t slot_a, int32_t slot_b);
int     ahfgpu_synchronize(ahfgpu_ctx *ctx);
int64_t ahfgpu_stage_count(ahfgpu_ctx *ctx, const char *name);   /* launches / items attributed to the stage */

#ifdef __cplusplus
}
#endif
#endif
