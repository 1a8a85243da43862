Module dlp_gpu_binding

  !!-----------------------------------------------------------------------
  !!
  !! ISO_C_BINDING module over libdlpgpu.so (include/dlpgpu.h): the B200 short-range two-body path as a drop-in
  !! for the existing DL_POLY 5.1.0 call sites
  !!
  !!   drivers.F90:675-679     Call link_cell_pairs(...)           -> Call link_cell_pairs_gpu(...)
  !!   two_body.F90:339-525    Do i = 1, config%natms  (vdw+ewald) \
  !!   two_body.F90:552-606    Do i = 1, config%natms  (excluded)   > Call two_body_pairs_gpu(...)
  !!   neighbours.F90:157-171  displacement loop of vnl_check      -> dlp_gpu_vnl_tolerance(...)
  !!
  !! Everything else of two_body_forces (SPME reciprocal space, gsum, long-range corrections, totals) stays Fortran.
  !! Pattern follows the OpenKIM coupling (kim.F90:12-19, :880-999).  INTEGRATION.md shows the edits to the call sites.
  !!
  !! This file is not compiled in the build image (no Fortran compiler there); it is written against the reference's
  !! types as of 5.1.0 and against the C prototypes in include/dlpgpu.h.
  !!
  !!-----------------------------------------------------------------------

  Use, Intrinsic :: iso_c_binding, Only: c_int, c_double, c_ptr, c_null_ptr, c_loc, c_char, c_associated, c_f_pointer, &
                                         c_long_long
  Use kinds,           Only: wi, wp
  Use comms,           Only: comms_type
  Use configuration,   Only: configuration_type
  Use constants,       Only: r4pie0
  Use domains,         Only: domains_type
  Use rdfs,            Only: rdf_type
  Use electrostatic,   Only: electrostatic_type, ELECTROSTATIC_COULOMB, ELECTROSTATIC_DDDP, &
                             ELECTROSTATIC_COULOMB_FORCE_SHIFT, ELECTROSTATIC_COULOMB_REACTION_FIELD
  Use errors_warnings, Only: error
  Use ewald,           Only: ewald_type
  Use neighbours,      Only: neighbours_type
  Use particle,        Only: corePart
  Use statistics,      Only: stats_type
  Use vdw,             Only: vdw_type

  Implicit None
  Private

  Type(c_ptr), Save :: ctx = c_null_ptr

  Public :: dlp_gpu_init, dlp_gpu_finalise, dlp_gpu_set_forcefield, link_cell_pairs_gpu, two_body_pairs_gpu, rdf_collect_gpu, &
            dlp_gpu_vnl_tolerance, dlp_gpu_set_host_threads

  Interface
    Function dlpgpu_create(ctx, device) Bind(C, name='dlpgpu_create') Result(rc)
      Import :: c_ptr, c_int
      Type(c_ptr), Intent(Out)      :: ctx
      Integer(c_int), Value         :: device
      Integer(c_int)                :: rc
    End Function
    Function dlpgpu_destroy(ctx) Bind(C, name='dlpgpu_destroy') Result(rc)
      Import :: c_ptr, c_int
      Type(c_ptr), Value :: ctx
      Integer(c_int)     :: rc
    End Function
    Function dlpgpu_last_error(ctx) Bind(C, name='dlpgpu_last_error') Result(msg)
      Import :: c_ptr
      Type(c_ptr), Value :: ctx
      Type(c_ptr)        :: msg
    End Function
    Function dlpgpu_set_domain(ctx, dd) Bind(C, name='dlpgpu_set_domain') Result(rc)
      Import :: c_ptr, c_int
      Type(c_ptr), Value         :: ctx
      Integer(c_int), Intent(In) :: dd(6)
      Integer(c_int)             :: rc
    End Function
    Function dlpgpu_set_cell(ctx, cell, imcon) Bind(C, name='dlpgpu_set_cell') Result(rc)
      Import :: c_ptr, c_int, c_double
      Type(c_ptr), Value          :: ctx
      Real(c_double), Intent(In)  :: cell(9)
      Integer(c_int), Value       :: imcon
      Integer(c_int)              :: rc
    End Function
    Function dlpgpu_set_cutoffs(ctx, rcut, padding, pdplnc) Bind(C, name='dlpgpu_set_cutoffs') Result(rc)
      Import :: c_ptr, c_int, c_double
      Type(c_ptr), Value    :: ctx
      Real(c_double), Value :: rcut, padding, pdplnc
      Integer(c_int)        :: rc
    End Function
    Function dlpgpu_set_vdw(ctx, ntypes, vdw_list, max_vdw, n_vdw, ltp, max_grid, tab_potential, tab_force, rvdw, &
                            force_shift, direct, param, afs, bfs) Bind(C, name='dlpgpu_set_vdw') Result(rc)
      Import :: c_ptr, c_int, c_double
      Type(c_ptr), Value    :: ctx
      Integer(c_int), Value :: ntypes, max_vdw, n_vdw, max_grid, force_shift, direct
      Type(c_ptr), Value    :: vdw_list, ltp, tab_potential, tab_force, param, afs, bfs
      Real(c_double), Value :: rvdw
      Integer(c_int)        :: rc
    End Function
    Function dlpgpu_set_ewald(ctx, active, alpha, scaling, nsamples, erfc_tab, erfc_deriv_tab, recip_spacing) &
      Bind(C, name='dlpgpu_set_ewald') Result(rc)
      Import :: c_ptr, c_int, c_double
      Type(c_ptr), Value    :: ctx
      Integer(c_int), Value :: active, nsamples
      Real(c_double), Value :: alpha, scaling, recip_spacing
      Type(c_ptr), Value    :: erfc_tab, erfc_deriv_tab
      Integer(c_int)        :: rc
    End Function
    Function dlpgpu_set_coulomb(ctx, kind, damp, scaling, force_shift, energy_shift, reaction_field, nsamples, erfc_tab, &
                                erfc_deriv_tab, recip_spacing) Bind(C, name='dlpgpu_set_coulomb') Result(rc)
      Import :: c_ptr, c_int, c_double
      Type(c_ptr), Value    :: ctx
      Integer(c_int), Value :: kind, damp, nsamples
      Real(c_double), Value :: scaling, force_shift, energy_shift, recip_spacing
      Real(c_double)        :: reaction_field(3)
      Type(c_ptr), Value    :: erfc_tab, erfc_deriv_tab
      Integer(c_int)        :: rc
    End Function
    Function dlpgpu_link_cell_pairs(ctx, natms, nlast, parts, ltype, ltg, lfrzn, lbook, megfrz, max_exclude, &
                                    list_excl, max_list, list_out, ibig) Bind(C, name='dlpgpu_link_cell_pairs') Result(rc)
      Import :: c_ptr, c_int
      Type(c_ptr), Value          :: ctx
      Integer(c_int), Value       :: natms, nlast, lbook, megfrz, max_exclude, max_list
      Type(c_ptr), Value          :: parts, ltype, ltg, lfrzn, list_excl, list_out
      Integer(c_int), Intent(Out) :: ibig
      Integer(c_int)              :: rc
    End Function
    Function dlpgpu_two_body_forces(ctx, natms, nlast, parts, out) Bind(C, name='dlpgpu_two_body_forces') Result(rc)
      Import :: c_ptr, c_int, c_double
      Type(c_ptr), Value           :: ctx
      Integer(c_int), Value        :: natms, nlast
      Type(c_ptr), Value           :: parts
      Real(c_double), Intent(Out)  :: out(16)
      Integer(c_int)               :: rc
    End Function
    Function dlpgpu_rdf_collect(ctx, ntypes, rdf_list, n_pairs, max_grid, rdf) Bind(C, name='dlpgpu_rdf_collect') Result(rc)
      Import :: c_ptr, c_int
      Type(c_ptr), Value    :: ctx
      Integer(c_int), Value :: ntypes, n_pairs, max_grid
      Type(c_ptr), Value    :: rdf_list, rdf
      Integer(c_int)        :: rc
    End Function
    Function dlpgpu_set_spme(ctx, kdim, nsplines) Bind(C, name='dlpgpu_set_spme') Result(rc)
      Import :: c_ptr, c_int
      Type(c_ptr), Value    :: ctx
      Integer(c_int)        :: kdim(3)
      Integer(c_int), Value :: nsplines
      Integer(c_int)        :: rc
    End Function
    Function dlpgpu_spme_forces(ctx, natms, parts, megatm, out) Bind(C, name='dlpgpu_spme_forces') Result(rc)
      Import :: c_ptr, c_int, c_double
      Type(c_ptr), Value          :: ctx
      Integer(c_int), Value       :: natms, megatm
      Type(c_ptr), Value          :: parts
      Real(c_double), Intent(Out) :: out(16)
      Integer(c_int)              :: rc
    End Function
    Function dlpgpu_set_host_threads(ctx, nthreads) Bind(C, name='dlpgpu_set_host_threads') Result(rc)
      Import :: c_ptr, c_int
      Type(c_ptr), Value    :: ctx
      Integer(c_int), Value :: nthreads
      Integer(c_int)        :: rc
    End Function
    Function dlpgpu_dev_spme_spread(ctx, grid_dev) Bind(C, name='dlpgpu_dev_spme_spread') Result(rc)
      Import :: c_ptr, c_int
      Type(c_ptr), Value :: ctx
      Type(c_ptr), Value :: grid_dev
      Integer(c_int)     :: rc
    End Function
    Function dlpgpu_dev_spme_solve_gather(ctx, grid_dev, ftot_local) Bind(C, name='dlpgpu_dev_spme_solve_gather') Result(rc)
      Import :: c_ptr, c_int, c_double
      Type(c_ptr), Value          :: ctx
      Type(c_ptr), Value          :: grid_dev
      Real(c_double), Intent(Out) :: ftot_local(3)
      Integer(c_int)              :: rc
    End Function
    Function dlpgpu_dev_spme_finish(ctx, megatm, ftot_global, nranks, out) Bind(C, name='dlpgpu_dev_spme_finish') Result(rc)
      Import :: c_ptr, c_int, c_double
      Type(c_ptr), Value          :: ctx
      Integer(c_int), Value       :: megatm, nranks
      Real(c_double), Intent(In)  :: ftot_global(3)
      Real(c_double), Intent(Out) :: out(16)
      Integer(c_int)              :: rc
    End Function
    Function dlpgpu_set_collect_pp(ctx, on) Bind(C, name='dlpgpu_set_collect_pp') Result(rc)
      Import :: c_ptr, c_int
      Type(c_ptr), Value    :: ctx
      Integer(c_int), Value :: on
      Integer(c_int)        :: rc
    End Function
    Function dlpgpu_get_pp(ctx, natms, pp_energy, pp_stress) Bind(C, name='dlpgpu_get_pp') Result(rc)
      Import :: c_ptr, c_int
      Type(c_ptr), Value    :: ctx
      Integer(c_int), Value :: natms
      Type(c_ptr), Value    :: pp_energy, pp_stress
      Integer(c_int)        :: rc
    End Function
    Function dlpgpu_parts_unchanged_since_list(ctx) Bind(C, name='dlpgpu_parts_unchanged_since_list') Result(rc)
      Import :: c_ptr, c_int
      Type(c_ptr), Value :: ctx
      Integer(c_int)     :: rc
    End Function
    Function dlpgpu_vnl_check(ctx, natms, parts, tol) Bind(C, name='dlpgpu_vnl_check') Result(rc)
      Import :: c_ptr, c_int, c_double
      Type(c_ptr), Value           :: ctx
      Integer(c_int), Value        :: natms
      Type(c_ptr), Value           :: parts
      Real(c_double), Intent(Out)  :: tol
      Integer(c_int)               :: rc
    End Function
    Function dlpgpu_vnl_set_check(ctx, nlast, parts) Bind(C, name='dlpgpu_vnl_set_check') Result(rc)
      Import :: c_ptr, c_int
      Type(c_ptr), Value    :: ctx
      Integer(c_int), Value :: nlast
      Type(c_ptr), Value    :: parts
      Integer(c_int)        :: rc
    End Function
  End Interface

Contains

  Subroutine check(rc, where)
    !! non-zero return codes carry DL_POLY's own error number where one exists (106, 95, 307, ...): abort all ranks the
    !! reference way (errors_warnings.F90:840)
    Integer(c_int),   Intent(In) :: rc
    Character(Len=*), Intent(In) :: where

    Character(Kind=c_char), Pointer :: cmsg(:)
    Character(Len=512)              :: msg
    Integer                         :: i
    Type(c_ptr)                     :: p

    If (rc == 0) Return
    msg = ' '
    p = dlpgpu_last_error(ctx)
    If (c_associated(p)) Then
      Call c_f_pointer(p, cmsg, [512])
      Do i = 1, 512
        If (cmsg(i) == Char(0)) Exit
        msg(i:i) = cmsg(i)
      End Do
    End If
    Call error(0, 'dlpgpu ('//Trim(where)//'): '//Trim(msg))
  End Subroutine check

  Subroutine dlp_gpu_init(device, domain, config, neigh, comm)
    !! once, after set_bounds / read_config: one context per MPI rank, GPU = local rank modulo GPUs per node
    Integer,                  Intent(In) :: device
    Type(domains_type),       Intent(In) :: domain
    Type(configuration_type), Intent(In) :: config
    Type(neighbours_type),    Intent(In) :: neigh
    Type(comms_type),         Intent(In) :: comm

    Integer(c_int) :: dd(6)

    Call check(dlpgpu_create(ctx, Int(device, c_int)), 'create')
    dd = [domain%nx, domain%ny, domain%nz, domain%idx, domain%idy, domain%idz]
    Call check(dlpgpu_set_domain(ctx, dd), 'set_domain')
    Call check(dlpgpu_set_cell(ctx, config%cell, Int(config%imcon, c_int)), 'set_cell')
    Call check(dlpgpu_set_cutoffs(ctx, neigh%cutoff, neigh%padding, neigh%pdplnc), 'set_cutoffs')
  End Subroutine dlp_gpu_init

  Subroutine dlp_gpu_set_host_threads(nthreads)
    !! how config%parts travels (include/dlpgpu.h): 0 = whole corePart records by DMA (default); n >= 1 = n host threads of the
    !! library copy x, y, z up and ADD the returned forces into parts%f -- use the cores the rank's OpenMP team would have had
    !! (a dozen or more pay off); in that mode the list_just_built assertion of two_body_pairs_gpu no longer covers parts%f
    Integer, Intent(In) :: nthreads

    Call check(dlpgpu_set_host_threads(ctx, Int(nthreads, c_int)), 'set_host_threads')
  End Subroutine dlp_gpu_set_host_threads

  Subroutine dlp_gpu_set_forcefield(ntpatm, vdws, electro, ewld, rcut, eps)
    !! once after vdw_generate / vdw_table_read / erfcgen (two_body.F90:188 computes the same scaling)
    Integer,                  Intent(In)         :: ntpatm
    Type(vdw_type),           Intent(In), Target :: vdws
    Type(electrostatic_type), Intent(In), Target :: electro
    Type(ewald_type),         Intent(In)         :: ewld
    Real(Kind=wp),            Intent(In)         :: rcut, eps

    Real(Kind=wp)  :: recip_spacing
    Type(c_ptr)    :: p_pot, p_frc, p_par, p_afs, p_bfs
    Integer(c_int) :: fs, dr, kind

    p_pot = c_null_ptr; p_frc = c_null_ptr; p_par = c_null_ptr; p_afs = c_null_ptr; p_bfs = c_null_ptr
    If (Allocated(vdws%tab_potential)) p_pot = c_loc(vdws%tab_potential)   ! (0:max_grid, 1:max_vdw), column-major
    If (Allocated(vdws%tab_force))     p_frc = c_loc(vdws%tab_force)
    If (Allocated(vdws%param))         p_par = c_loc(vdws%param)           ! (1:max_param=7.., 1:max_vdw)
    If (Allocated(vdws%afs))           p_afs = c_loc(vdws%afs)
    If (Allocated(vdws%bfs))           p_bfs = c_loc(vdws%bfs)
    fs = Merge(1, 0, vdws%l_force_shift); dr = Merge(1, 0, vdws%l_direct)
    If (vdws%no_vdw .or. vdws%n_vdw <= 0) Then
      Call check(dlpgpu_set_vdw(ctx, Int(ntpatm, c_int), c_null_ptr, 0_c_int, 0_c_int, c_null_ptr, 0_c_int, c_null_ptr, &
                                c_null_ptr, 0.0_c_double, 0_c_int, 0_c_int, c_null_ptr, c_null_ptr, c_null_ptr), 'set_vdw')
    Else
      Call check(dlpgpu_set_vdw(ctx, Int(ntpatm, c_int), c_loc(vdws%list), Int(vdws%max_vdw, c_int), Int(vdws%n_vdw, c_int), &
                                c_loc(vdws%ltp), Int(vdws%max_grid, c_int), p_pot, p_frc, vdws%cutoff, fs, dr, p_par, p_afs, &
                                p_bfs), 'set_vdw')
    End If
    If (Any(electro%key == [ELECTROSTATIC_COULOMB, ELECTROSTATIC_DDDP, ELECTROSTATIC_COULOMB_FORCE_SHIFT, &
                            ELECTROSTATIC_COULOMB_REACTION_FIELD])) Then
      ! coul_spole.F90: the direct-space variants dispatched at two_body.F90:480-514.  force_shift / energy_shift /
      ! reaction_field are set by the reference's own first call (coul_spole.F90:186-202, :407-417); call this routine after
      ! them or reproduce those statements here.  Damped forms pass the erfc tables generated with electro%damping.
      Select Case (electro%key)
      Case (ELECTROSTATIC_COULOMB);                kind = 1
      Case (ELECTROSTATIC_DDDP);                   kind = 2
      Case (ELECTROSTATIC_COULOMB_FORCE_SHIFT);    kind = 3
      Case Default;                                kind = 4
      End Select
      If (electro%damp .and. kind >= 3) Then
        recip_spacing = 1.0_wp / (rcut / Real(electro%erfc%nsamples - 4, wp))
        Call check(dlpgpu_set_coulomb(ctx, kind, 1_c_int, r4pie0 / eps, electro%force_shift, electro%energy_shift, &
                                      electro%reaction_field(0:2), Int(electro%erfc%nsamples, c_int), &
                                      table_base(electro%erfc%table), table_base(electro%erfc_deriv%table), recip_spacing), &
                   'set_coulomb')
      Else
        Call check(dlpgpu_set_coulomb(ctx, kind, 0_c_int, r4pie0 / eps, electro%force_shift, electro%energy_shift, &
                                      electro%reaction_field(0:2), 0_c_int, c_null_ptr, c_null_ptr, 0.0_c_double), 'set_coulomb')
      End If
    Else If (electro%erfc%initialised) Then
      ! interp_table%recip_spacing is private: recompute it with the statements of init_interp_table (numerics.F90:237-238).
      ! The C side reads element i of the array as table(i); table(1:nsamples) has no element 0, so pass the address one
      ! element before table(1) -- it is never dereferenced for r >= spacing.
      recip_spacing = 1.0_wp / (rcut / Real(electro%erfc%nsamples - 4, wp))
      Call check(dlpgpu_set_ewald(ctx, 1_c_int, ewld%alpha, r4pie0 / eps, Int(electro%erfc%nsamples, c_int), &
                                  table_base(electro%erfc%table), table_base(electro%erfc_deriv%table), recip_spacing), &
                 'set_ewald')
    Else
      Call check(dlpgpu_set_ewald(ctx, 0_c_int, 0.0_c_double, 0.0_c_double, 0_c_int, c_null_ptr, c_null_ptr, 0.0_c_double), &
                 'set_ewald')
    End If
  End Subroutine dlp_gpu_set_forcefield

  Function table_base(table) Result(p)
    !! address of the (non-existent) element 0 of table(1:n): a copy shifted by one keeps the C indexing simple and safe
    Real(Kind=wp), Intent(In), Target :: table(:)
    Type(c_ptr)                       :: p

    Real(Kind=wp), Allocatable, Save, Target :: shifted(:, :)
    Integer, Save                            :: used = 0

    If (Allocated(shifted)) Then
      If (Ubound(shifted, 1) /= Size(table)) Deallocate (shifted)
    End If
    If (.not. Allocated(shifted)) Allocate (shifted(0:Size(table), 2))
    used = Mod(used, 2) + 1     ! two tables per call of dlp_gpu_set_forcefield; the library copies them before returning
    shifted(0, used) = 0.0_wp
    shifted(1:Size(table), used) = table(:)
    p = c_loc(shifted(0, used))
  End Function table_base

  Subroutine link_cell_pairs_gpu(lbook, megfrz, neigh, config, want_host_list)
    !! replaces Call link_cell_pairs(...) (neighbours.F90:356).  The list stays on the device; neigh%list is filled only
    !! when another consumer needs it (rdf, metal, kim, ...: want_host_list).  Also takes the vnl_set_check snapshot.
    Logical,                  Intent(In   )         :: lbook
    Integer,                  Intent(In   )         :: megfrz
    Type(neighbours_type),    Intent(InOut), Target :: neigh
    Type(configuration_type), Intent(InOut), Target :: config
    Logical,                  Intent(In   )         :: want_host_list

    Integer(c_int) :: ibig, rc
    Type(c_ptr)    :: p_list, p_excl

    p_list = c_null_ptr; p_excl = c_null_ptr
    If (want_host_list) p_list = c_loc(neigh%list)                    ! (-3:max_list, 1:mxatdm)
    If (lbook) p_excl = c_loc(neigh%list_excl)                        ! (0:max_exclude, 1:mxatdm)
    rc = dlpgpu_link_cell_pairs(ctx, Int(config%natms, c_int), Int(config%nlast, c_int), c_loc(config%parts), &
                                c_loc(config%ltype), c_loc(config%ltg), c_loc(config%lfrzn), Merge(1_c_int, 0_c_int, lbook), &
                                Int(megfrz, c_int), Int(neigh%max_exclude, c_int), p_excl, Int(neigh%max_list, c_int), &
                                p_list, ibig)
    If (rc == 106) Then
      Call error(106, 'neighbour list array exceeded; largest row', .true.)   ! neighbours.F90:1189-1194 reports ibig
    End If
    Call check(rc, 'link_cell_pairs')
  End Subroutine link_cell_pairs_gpu

  Subroutine two_body_pairs_gpu(config, stats, engvdw, virvdw, engcpe_rl, vircpe_rl, engcpe_ex, vircpe_ex, list_just_built)
    !! replaces the two Do i = 1, config%natms loops of two_body_forces: forces are ADDED into config%parts(1:natms)%f,
    !! the six per-rank partial sums feed the existing gsum buffer (two_body.F90:708-729) and the nine stress
    !! contributions are added to stats%stress exactly where vdw_forces / ewald_real_forces add theirs.
    Type(configuration_type), Intent(InOut), Target :: config
    Type(stats_type),         Intent(InOut)         :: stats
    Real(Kind=wp),            Intent(InOut)         :: engvdw, virvdw, engcpe_rl, vircpe_rl, engcpe_ex, vircpe_ex
    Logical, Optional,        Intent(In   )         :: list_just_built   !! link_cell_pairs_gpu ran in this calculate_forces
                                                                         !! AND nothing has written config%parts since --
                                                                         !! forces included unless dlpgpu_set_host_threads
                                                                         !! chose the packed mode: pass neigh%update only
                                                                         !! when no tersoff / three-body / four-body
                                                                         !! provider is active (they add into parts%f
                                                                         !! between the two calls, drivers.F90:675-700).

    Real(c_double) :: out(16)

    ! per-particle energy / stress (statistics.F90:227): the library books them inside the same pair loop
    Call check(dlpgpu_set_collect_pp(ctx, Merge(1_c_int, 0_c_int, stats%collect_pp)), 'set_collect_pp')

    If (Present(list_just_built)) Then
      If (list_just_built) Call check(dlpgpu_parts_unchanged_since_list(ctx), 'parts_unchanged_since_list')
    End If

    Call check(dlpgpu_two_body_forces(ctx, Int(config%natms, c_int), Int(config%nlast, c_int), c_loc(config%parts), out), &
               'two_body_forces')
    engvdw = engvdw + out(1); virvdw = virvdw + out(2)
    engcpe_rl = engcpe_rl + out(3); vircpe_rl = vircpe_rl + out(4)
    engcpe_ex = engcpe_ex + out(5); vircpe_ex = vircpe_ex + out(6)
    stats%stress(1:9) = stats%stress(1:9) + out(7:15)
    If (stats%collect_pp) Then   ! vdw.F90:1741-1755, :1987-2001, ewald_spole.F90:205-215: added to what other providers booked
      Call check(dlpgpu_get_pp(ctx, Int(config%natms, c_int), c_loc(stats%pp_energy), c_loc(stats%pp_stress)), 'get_pp')
    End If
  End Subroutine two_body_pairs_gpu

  Subroutine ewald_spme_forces_gpu(ewld, config, stats, engcpe_rc, vircpe_rc)
    !! one-domain counterpart of ewald_spme_forces_coul (ewald_spole.F90:244-477; call site two_body.F90:298-302): reciprocal forces
    !! are ADDED into config%parts(1:natms)%f, the stress contributions into stats%stress.  comm%mxnode == 1 only; call it before
    !! link_cell_pairs_gpu of the step (it reuses the device atom arrays).
    Type(ewald_type),         Intent(In   )         :: ewld
    Type(configuration_type), Intent(InOut), Target :: config
    Type(stats_type),         Intent(InOut)         :: stats
    Real(Kind=wp),            Intent(  Out)         :: engcpe_rc, vircpe_rc

    Real(c_double), Save :: out(16)
    Logical,        Save :: grid_set = .false.
    Integer(c_int)       :: kdim(3)

    If (.not. grid_set) Then
      kdim = Int(ewld%kspace%k_vec_dim, c_int)
      Call check(dlpgpu_set_spme(ctx, kdim, Int(ewld%bspline%num_splines, c_int)), 'set_spme')
      grid_set = .true.
    End If
    Call check(dlpgpu_spme_forces(ctx, Int(config%natms, c_int), c_loc(config%parts), Int(config%megatm, c_int), out), 'spme_forces')
    engcpe_rc = out(1); vircpe_rc = out(2)
    stats%stress(1:9) = stats%stress(1:9) + out(3:11)
  End Subroutine ewald_spme_forces_gpu

  Subroutine rdf_collect_gpu(ntpatm, rdf)
    !! replaces the per-atom Call rdf_collect / rdf_excl_collect inside two_body_forces (two_body.F90:523, :581) on steps with
    !! l_do_rdf: the counts of this rank's pairs are added to rdf%rdf(1:max_grid, 1:n_pairs).  Call after two_body_pairs_gpu.
    !! (rdf%tmp_rdf block statistics and rdf_frzn_collect stay on the host path.)
    Integer,        Intent(In   )         :: ntpatm
    Type(rdf_type), Intent(InOut), Target :: rdf

    Call check(dlpgpu_rdf_collect(ctx, Int(ntpatm, c_int), c_loc(rdf%list), Int(rdf%n_pairs, c_int), Int(rdf%max_grid, c_int), &
                                  c_loc(rdf%rdf)), 'rdf_collect')
  End Subroutine rdf_collect_gpu

  Function dlp_gpu_vnl_tolerance(config) Result(tol)
    !! neighbours.F90:157-171: max_i |r_i - r_bg,i| over the local atoms; the caller keeps gmax and the comparison
    Type(configuration_type), Intent(In), Target :: config
    Real(Kind=wp)                                :: tol

    Call check(dlpgpu_vnl_check(ctx, Int(config%natms, c_int), c_loc(config%parts), tol), 'vnl_check')
  End Function dlp_gpu_vnl_tolerance

  Subroutine dlp_gpu_finalise()
    If (c_associated(ctx)) Call check(dlpgpu_destroy(ctx), 'destroy')
    ctx = c_null_ptr
  End Subroutine dlp_gpu_finalise

End Module dlp_gpu_binding
