# TTNEvalB200.jl — the Julia side of the drop-in: a batched `evaluate` method and the packer that
# feeds libttneval.so through `ccall`.
#
# Where it goes in the reference: a new file `src/ttneval_b200.jl`, included from
# `src/ITensorNumericalAnalysis.jl` right after `itensornetworkfunction.jl` (it only needs
# `ITensorNetworkFunction`, `indsnetworkmap`, `indexmap`, `index_value_to_scalar`, `dimension`,
# `digit`, `is_real`).  It adds methods, it replaces nothing: the scalar-point methods
# (src/itensornetworkfunction.jl:96-112) keep working unchanged, and
#     evaluate(fitn, points::Vector{<:Vector}, dims)      # new: many points
#     evaluate(fitn, points::AbstractMatrix, dims)        # new: D x Npts matrix
# are strictly more specific than `xs::Vector`, so there is no dispatch ambiguity (the same trick
# as delta_p, src/elementary_functions.jl:224-229 vs 249-254).
#
# STATUS: written against ITensorNetworks 0.13 / ITensors 0.9 / NamedGraphs 0.6 as used by the
# reference; NOT executed in the build environment (no Julia in the image).  The Python mirror
# (../packer.py, ../itensornetworkfunction.py) is the executed twin of this file; both produce
# the same `ttn_desc` (include/ttneval.h).

using ITensors: ITensors, Index, dim, array, permute, inds, commoninds, hastags
using ITensorNetworks: ITensorNetworks, ITensorNetwork, siteinds
using Graphs: Graphs, vertices, neighbors, edges, nv
using NamedGraphs.GraphsExtensions: is_tree, leaf_vertices

const LIBTTNEVAL = get(ENV, "LIBTTNEVAL", "libttneval.so")

const TTN_ABI_VERSION = Int32(3)
const TTN_LAYOUT_AOS = Int32(0)   # coords[c + n_coords*p]: a Julia (n_coords x npts) Matrix
const TTN_MEM_HOST = Int32(0)

# mirror of `struct ttn_desc` (include/ttneval.h)
struct TTNDesc
  abi_version::Int32
  n_vertices::Int32
  n_coords::Int32
  is_complex::Int32
  root::Int32
  n_sites::Int32
  parent::Ptr{Int32}
  link_dim::Ptr{Int32}
  site_ptr::Ptr{Int32}
  site_dim::Ptr{Int32}
  site_coord::Ptr{Int32}
  site_digit::Ptr{Int32}
  thr_ptr::Ptr{Int32}
  thr::Ptr{Float64}
  tensor_ptr::Ptr{Int64}
  tensors::Ptr{Cvoid}
end

# mirror of `struct ttn_opts`
mutable struct TTNOpts
  coords_mem::Int32
  out_mem::Int32
  kernel::Int32
  reduce_sum::Int32
  chunk_points::Int64
  sum_re::Float64
  sum_im::Float64
  kernel_ms::Float32
  total_ms::Float32
  kernel_used::Int32
  n_launches::Int32
  weights::Ptr{Float64}
  weights_mem::Int32
  reserved_::Int32
  flops_executed::Float64
  host_staging::Int32   # 0 = pageable Julia arrays go through the library's pinned staging ring (run-path coordinates quantised
                        # to their grid index on the way), 1 = raw cudaMemcpyAsync, 2 = ring, coordinates stay doubles
  accuracy::Int32       # 0 = FP64, 1 = refined (double-double re-evaluation of the points that cancel)
  refine_tau::Float64
  n_devices_used::Int32
  staged::Int32
  n_refined::Int64
  h2d_bytes::Int64
  d2h_bytes::Int64
end
# reduce: TTN_REDUCE_NONE = 0, _SUM = 1 (sum f), _ABS2 = 2 (sum |f|^2), _WEIGHTED = 3 (sum w f)
const REDUCE_MODES = Dict(:none => Int32(0), :sum => Int32(1), :abs2 => Int32(2), :weighted => Int32(3))
const ACCURACY_MODES = Dict(:fp64 => Int32(0), :refined => Int32(1))
TTNOpts(; reduce::Symbol=:none, weights::Ptr{Float64}=Ptr{Float64}(C_NULL), accuracy::Symbol=:fp64) =
  TTNOpts(TTN_MEM_HOST, TTN_MEM_HOST, 0, REDUCE_MODES[reduce], 0, 0.0, 0.0, 0.0f0, 0.0f0, 0, 0, weights, TTN_MEM_HOST, 0, 0.0,
    0, ACCURACY_MODES[accuracy], 0.0, 0, 0, 0, 0, 0)

"Flat arrays of one packed network; keeps everything the C side points at alive."
struct PackedNetwork
  parent::Vector{Int32}
  link_dim::Vector{Int32}
  site_ptr::Vector{Int32}
  site_dim::Vector{Int32}
  site_coord::Vector{Int32}
  site_digit::Vector{Int32}
  thr_ptr::Vector{Int32}
  thr::Vector{Float64}
  tensor_ptr::Vector{Int64}
  tensors::Vector            # Vector{Float64} or Vector{ComplexF64}
  root::Int32
  n_coords::Int32
  is_complex::Bool
  complex_coords::Bool
  site_inds::Vector{Index}   # site index of column s of the description (evaluate_indices, ind_values)
end

mutable struct TTNPlan
  handle::Ptr{Cvoid}
  packed::PackedNetwork
  # ngpus > 1: the plan is replicated on GPUs 0..ngpus-1 of this process and every call shards its points over
  # them (ttn_plan_create_multi): contiguous blocks, each GPU copies its block in and its values out itself
  function TTNPlan(packed::PackedNetwork; device::Integer=0, ngpus::Integer=1)
    desc = TTNDesc(
      TTN_ABI_VERSION, length(packed.parent), packed.n_coords, packed.is_complex ? 1 : 0,
      packed.root, length(packed.site_dim),
      pointer(packed.parent), pointer(packed.link_dim), pointer(packed.site_ptr),
      pointer(packed.site_dim), pointer(packed.site_coord), pointer(packed.site_digit),
      pointer(packed.thr_ptr), pointer(packed.thr), pointer(packed.tensor_ptr),
      Ptr{Cvoid}(pointer(packed.tensors)),
    )
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve packed begin
      rc = if ngpus > 1
        ccall((:ttn_plan_create_multi, LIBTTNEVAL), Cint, (Ref{TTNDesc}, Int32, Ptr{Int32}, Ref{Ptr{Cvoid}}),
          desc, Int32(ngpus), C_NULL, h)
      else
        ccall((:ttn_plan_create, LIBTTNEVAL), Cint, (Ref{TTNDesc}, Int32, Ref{Ptr{Cvoid}}),
          desc, Int32(device), h)
      end
    end
    rc == 0 || error("ttn_plan_create: " * unsafe_string(ccall((:ttn_last_error, LIBTTNEVAL), Cstring, ())))
    plan = new(h[], packed)
    finalizer(p -> ccall((:ttn_plan_destroy, LIBTTNEVAL), Cvoid, (Ptr{Cvoid},), p.handle), plan)
    return plan
  end
end

"Root choice: an end vertex for path graphs (MPS); the most central vertex of degree <= 2 for trees of maximum degree 3; otherwise a centre of the tree."
function choose_root(g, vs)
  length(vs) == 1 && return first(vs)
  if maximum(v -> length(neighbors(g, v)), vs) <= 2
    return last(filter(v -> length(neighbors(g, v)) == 1, vs))
  end
  deg = Dict(v => length(neighbors(g, v)) for v in vs)
  if maximum(values(deg)) == 3
    # rooted at a vertex of degree <= 2 no vertex has more than two children (per-vertex GEMM kernel);
    # take the most central such vertex
    function ecc(r)
      dist = Dict(r => 0)
      todo = [r]
      i = 1
      while i <= length(todo)
        v = todo[i]; i += 1
        for u in neighbors(g, v)
          haskey(dist, u) || (dist[u] = dist[v] + 1; push!(todo, u))
        end
      end
      return maximum(values(dist))
    end
    cands = filter(v -> deg[v] <= 2, vs)
    return cands[argmin([(ecc(v), findfirst(==(v), vs)) for v in cands])]
  end
  remaining = Set(vs)
  leaves = filter(v -> deg[v] == 1, vs)
  while length(remaining) > 2
    nxt = eltype(vs)[]
    for v in leaves
      delete!(remaining, v)
      for u in neighbors(g, v)
        if u in remaining
          deg[u] -= 1
          deg[u] == 1 && push!(nxt, u)
        end
      end
    end
    leaves = nxt
  end
  return first(filter(v -> v in remaining, vs))
end

"""
    pack(fitn, dims) -> PackedNetwork

Runs once per (network, dims).  Replaces the per-point `copy(fitn)`
(src/itensornetworkfunction.jl:85), dictionary filters / `sort`
(src/IndexMaps/realindexmap.jl:67-76) and `project` bookkeeping
(src/itensornetworkfunction.jl:84-94).  Thresholds come from the reference's own
`index_value_to_scalar`, so digit selection is bit-identical for every base.
"""
function pack(fitn::ITensorNetworkFunction, dims::Vector{<:Int}=dimensions(fitn))
  tn = itensornetwork(fitn)
  # Loopy graphs: the reference only evaluates them with link dimension 1 (const_itn on named_grid((3,3)),
  # test/test_realitensorfunction.jl:39-57) or alg="exact".  A dimension-1 link carries no sum, so the links
  # outside a spanning tree are dropped exactly (their axes have length 1); anything else stays on the
  # reference's per-point path.
  if !is_tree(fitn)
    all(e -> dim(only(commoninds(tn[Graphs.src(e)], tn[Graphs.dst(e)]))) == 1, edges(tn)) ||
      error("the batched evaluator needs a tree, or a loopy network whose link dimensions are all 1 (cf. truncate, src/itensornetworkfunction.jl:115)")
  end
  s = indsnetwork(indsnetworkmap(fitn))
  imap = indexmap(fitn)
  cmap = imap isa ComplexIndexMap
  all(d -> d in dims, dimensions(imap)) ||
    throw(KeyError("dims $dims do not cover all dimensions $(dimensions(imap)) of the network"))
  vs = collect(vertices(tn))
  vid = Dict(v => Int32(i - 1) for (i, v) in enumerate(vs))
  root = choose_root(tn, vs)
  parent = fill(Int32(-1), length(vs))
  order = [root]
  seen = Set([root])
  for v in order, u in neighbors(tn, v)
    if !(u in seen)
      push!(seen, u); push!(order, u)
      parent[vid[u] + 1] = vid[v]
    end
  end
  eltype_c = any(v -> eltype(tn[v]) <: Complex, vs)
  T = eltype_c ? ComplexF64 : Float64
  link_dim = ones(Int32, length(vs))
  site_ptr = Int32[0]; site_dim = Int32[]; site_coord = Int32[]; site_digit = Int32[]
  thr_ptr = Int32[0]; thr = Float64[]
  tensor_ptr = Int64[0]; tensors = T[]
  site_inds = Index[]
  for v in vs
    i = vid[v]
    sites = collect(s[v])
    children = sort(filter(u -> parent[vid[u] + 1] == i, collect(neighbors(tn, v))); by=u -> vid[u])
    # neighbours that are neither children nor the parent: dimension-1 links outside the BFS spanning tree
    extra = Index[only(commoninds(tn[v], tn[u])) for u in neighbors(tn, v)
                  if parent[vid[u] + 1] != i && parent[i + 1] != vid[u]]
    # C order [site..., child..., parent] (last fastest) == Julia column-major with reversed axes
    axes = Index[]
    append!(axes, sites)
    for c in children
      push!(axes, only(commoninds(tn[v], tn[c])))
    end
    if parent[i + 1] >= 0
      pl = only(commoninds(tn[v], tn[vs[parent[i + 1] + 1]]))
      push!(axes, pl)
      link_dim[i + 1] = dim(pl)
    end
    arr = (isempty(axes) && isempty(extra)) ? T[tn[v][]] : vec(array(permute(tn[v], reverse(axes)..., extra...)))
    append!(tensors, T.(arr))
    push!(tensor_ptr, length(tensors))
    for ind in sites
      pos = findfirst(==(dimension(imap, ind)), dims) - 1
      slot = cmap ? 2 * pos + (is_real(imap, ind) ? 0 : 1) : pos
      push!(site_dim, dim(ind)); push!(site_coord, slot); push!(site_digit, digit(imap, ind))
      push!(site_inds, ind)
      append!(thr, [abs(index_value_to_scalar(imap, ind, k)) for k in 0:(dim(ind) - 1)])
      push!(thr_ptr, length(thr))
    end
    push!(site_ptr, length(site_dim))
  end
  return PackedNetwork(parent, link_dim, site_ptr, site_dim, site_coord, site_digit, thr_ptr, thr,
    tensor_ptr, tensors, vid[root], (cmap ? 2 : 1) * length(dims), eltype_c, cmap, site_inds)
end

# ---- plan cache -------------------------------------------------------------------------------------------
# Keyed on the ITensorNetwork OBJECT (a mutable struct) in a WeakKeyDict, so a network that is garbage collected
# takes its plans (up to ~0.5 GB of device memory each) with it: their finalizers run ttn_plan_destroy.  The
# reference mutates networks in place (`psi[v] = t`, `psi[v] *= c`), and both create a NEW ITensor — a new storage
# object — at that vertex: every entry therefore carries a fingerprint of the tensors it was packed from (the
# objectid of every vertex tensor's storage plus one sampled element), and a plan whose fingerprint no longer
# matches is rebuilt.  Not covered: writing single ELEMENTS into a tensor's storage (`psi[v][i, j] = x`) other
# than the sampled one — call `invalidate_plans!(fitn)` after that, or pass an explicit plan:
#     pl = TTNPlan(pack(fitn, dims)); evaluate(pl, points)
const _plan_cache = WeakKeyDict{Any,Any}()
const _plan_cache_lock = ReentrantLock()

function network_fingerprint(fitn::ITensorNetworkFunction)
  tn = itensornetwork(fitn)
  h = UInt(0x7474)
  for v in vertices(tn)
    t = tn[v]
    st = ITensors.storage(t)
    h = hash((objectid(st), inds(t), isempty(ITensors.data(st)) ? 0.0 : first(ITensors.data(st))), h)
  end
  return h
end

function plan(fitn::ITensorNetworkFunction, dims; device=0, ngpus=1)
  fp = network_fingerprint(fitn)
  lock(_plan_cache_lock) do
    entry = get!(() -> Dict{Any,Any}(), _plan_cache, itensornetwork(fitn))
    key = (collect(dims), device, ngpus)
    hit = get(entry, key, nothing)
    if hit === nothing || hit[1] != fp
      entry[key] = (fp, TTNPlan(pack(fitn, dims); device, ngpus))
    end
    return entry[key][2]
  end
end

"Drop every cached plan of `fitn` (needed only after element-wise writes into a tensor's storage)."
invalidate_plans!(fitn::ITensorNetworkFunction) =
  lock(() -> (delete!(_plan_cache, itensornetwork(fitn)); nothing), _plan_cache_lock)

"coords as the (n_coords x npts) Float64 matrix the C side reads in AOS layout"
function coords_matrix(packed::PackedNetwork, points::AbstractMatrix)
  if packed.complex_coords
    z = ComplexF64.(points)
    out = Matrix{Float64}(undef, 2 * size(z, 1), size(z, 2))
    out[1:2:end, :] .= real.(z); out[2:2:end, :] .= imag.(z)
    return out
  end
  return Matrix{Float64}(points)
end

"""
    evaluate(fitn, points::AbstractMatrix, dims; reduce=:none, weights=nothing, device=0, ngpus=1, accuracy=:fp64)
    evaluate(fitn, points::Vector{<:Vector}, dims; ...)
    evaluate(plan::TTNPlan, points::AbstractMatrix; ...)        # explicit plan, no cache lookup

Batched `evaluate`: column `j` of `points` (or `points[j]`) holds the coordinates of point `j`
along `dims`.  Returns `Vector{Float64}` for real networks and `Vector{ComplexF64}` for complex
ones.  `reduce = :sum` returns the sum over all points, `:abs2` the sum of |f|^2, `:weighted` (with
`weights`, one real weight per point) the weighted sum — the fused quadrature functionals, computed in
the kernels' epilogues without writing the values.  Negative or NaN coordinates raise an error (the
reference's digit loop does not terminate on them).  `ngpus = G` shards the points over GPUs 0..G-1 of this
process (contiguous blocks, the network replicated, every GPU copies its own block in and its values out;
the per-GPU sums of a `reduce` are added in device order).  `accuracy = :refined` re-evaluates the points whose
value is small against the RMS of the batch (cancellation) in double-double arithmetic, so that the relative
error bound 1e-12 holds at the maximum and not only at the 99.9th percentile.  `points` may be any Julia
array: pageable memory is staged through the library's pinned ring by several host threads.
"""
function evaluate(fitn::ITensorNetworkFunction, points::AbstractMatrix,
  dims::Vector{<:Int}=dimensions(fitn); alg=default_contraction_alg(), device=0, ngpus=1, kwargs...)
  @assert size(points, 1) == length(dims)
  return evaluate(plan(fitn, dims; device, ngpus), points; kwargs...)
end

function evaluate(pl::TTNPlan, points::AbstractMatrix; reduce::Symbol=:none,
  weights::Union{Nothing,Vector{Float64}}=nothing, accuracy::Symbol=:fp64)
  coords = coords_matrix(pl.packed, points)
  npts = size(coords, 2)
  T = pl.packed.is_complex ? ComplexF64 : Float64
  out = reduce == :none ? Vector{T}(undef, npts) : T[]
  w = weights === nothing ? Float64[] : weights
  reduce == :weighted && @assert length(w) == npts
  opts = TTNOpts(; reduce, weights=(reduce == :weighted ? pointer(w) : Ptr{Float64}(C_NULL)), accuracy)
  GC.@preserve coords out pl w begin
    rc = ccall((:ttn_evaluate, LIBTTNEVAL), Cint,
      (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Int32, Ptr{Cvoid}, Ref{TTNOpts}),
      pl.handle, coords, npts, size(coords, 1), TTN_LAYOUT_AOS,
      reduce == :none ? pointer(out) : C_NULL, opts)
  end
  rc == 0 || error("ttn_evaluate: " * unsafe_string(ccall((:ttn_last_error, LIBTTNEVAL), Cstring, ())))
  reduce == :none && return out
  return (pl.packed.is_complex && reduce != :abs2) ? complex(opts.sum_re, opts.sum_im) : opts.sum_re
end

function evaluate(fitn::ITensorNetworkFunction, points::Vector{<:Vector},
  dims::Vector{<:Int}=dimensions(fitn); kwargs...)
  return evaluate(fitn, reduce(hcat, points), dims; kwargs...)
end

# mirror of `struct ttn_grid`
struct TTNGrid
  n_coords::Int32
  step::Ptr{Float64}
  count::Ptr{Int64}
  first::Int64
  npts::Int64
end

"""
    evaluate_grid(fitn, N::Int, dims=dimensions(fitn); reduce=:sum, values=false, device=0)

`fitn` on the Cartesian product of `grid_points(fitn, N, d)`, `d in dims`
(src/IndexMaps/realindexmap.jl:78-86; first dimension slowest), generated on the device: no coordinate
array exists on either side.  Returns the reduction (`reduce = :sum`: with `N = base^L` this is
`integrate(fitn; take_sum=true)`, src/integration.jl:6-17), the values (`values = true`), or both as a tuple.
The loop it replaces: examples/2d_laplace_solver.jl:46-53.  On the full dyadic grid of an MPS the
library shares the common digit prefixes of neighbouring grid points.
"""
function evaluate_grid(fitn::ITensorNetworkFunction, N::Int, dims::Vector{<:Int}=dimensions(fitn);
  reduce::Symbol=:sum, values::Bool=false, device=0, ngpus=1)
  imap = indexmap(fitn)
  @assert imap isa RealIndexMap "grid evaluation is defined for real index maps (grid_points)"
  pl = plan(fitn, dims; device, ngpus)
  grids = [grid_points(imap, N, d) for d in dims]
  steps = Float64[length(g) > 1 ? g[2] : 0.0 for g in grids]
  counts = Int64[length(g) for g in grids]
  npts = prod(counts)
  T = pl.packed.is_complex ? ComplexF64 : Float64
  out = values ? Vector{T}(undef, npts) : T[]
  opts = TTNOpts(; reduce)
  GC.@preserve steps counts out pl begin
    grid = TTNGrid(Int32(length(dims)), pointer(steps), pointer(counts), 0, npts)
    rc = ccall((:ttn_evaluate_grid, LIBTTNEVAL), Cint, (Ptr{Cvoid}, Ref{TTNGrid}, Ptr{Cvoid}, Ref{TTNOpts}),
      pl.handle, grid, values ? pointer(out) : C_NULL, opts)
  end
  rc == 0 || error("ttn_evaluate_grid: " * unsafe_string(ccall((:ttn_last_error, LIBTTNEVAL), Cstring, ())))
  red = (pl.packed.is_complex && reduce != :abs2) ? complex(opts.sum_re, opts.sum_im) : opts.sum_re
  return reduce == :none ? out : (values ? (red, out) : red)
end

"""
    evaluate_indices(fitn, ind_to_ind_value_maps::Vector; device=0)

Batched `project` + `scalar` (src/itensornetworkfunction.jl:84-106) at given index settings — what
`calculate_ind_values` returns, one dictionary per point — without the coordinate -> digit step: the inner
loop of TCI (ext/ITensorNumericalAnalysisTCIExt/tci_util.jl:21-55) evaluates its pivots and fibres this way.
"""
function evaluate_indices(fitn::ITensorNetworkFunction, ind_to_ind_value_maps::Vector; device=0)
  pl = plan(fitn, dimensions(fitn); device)
  sites = pl.packed.site_inds
  npts = length(ind_to_ind_value_maps)
  iv = Matrix{UInt8}(undef, length(sites), npts)       # column-major [site, point] == C [point][site]
  for (j, m) in enumerate(ind_to_ind_value_maps), (i, ind) in enumerate(sites)
    iv[i, j] = UInt8(m[ind])
  end
  T = pl.packed.is_complex ? ComplexF64 : Float64
  out = Vector{T}(undef, npts)
  opts = TTNOpts()
  GC.@preserve iv out pl begin
    rc = ccall((:ttn_evaluate_indices, LIBTTNEVAL), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64, Ptr{Cvoid}, Ref{TTNOpts}),
      pl.handle, iv, npts, pointer(out), opts)
  end
  rc == 0 || error("ttn_evaluate_indices: " * unsafe_string(ccall((:ttn_last_error, LIBTTNEVAL), Cstring, ())))
  return out
end

"""
    ind_values(fitn, points::AbstractMatrix, dims=dimensions(fitn)) -> (Matrix{UInt8}, Vector{Index})

Batched `calculate_ind_values` (src/IndexMaps/realindexmap.jl:67-76, complexindexmap.jl:116-132): entry
`[s, j]` is the value chosen for site index `site_inds[s]` at point `j`; bit-identical to the reference's
greedy loop (the thresholds come from the reference's own `index_value_to_scalar`).
"""
function ind_values(fitn::ITensorNetworkFunction, points::AbstractMatrix, dims::Vector{<:Int}=dimensions(fitn); device=0)
  pl = plan(fitn, dims; device)
  coords = coords_matrix(pl.packed, points)
  npts = size(coords, 2)
  dig = Matrix{UInt8}(undef, length(pl.packed.site_inds), npts)
  opts = TTNOpts()
  GC.@preserve coords dig pl begin
    rc = ccall((:ttn_digits, LIBTTNEVAL), Cint,
      (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Int32, Ptr{UInt8}, Ref{TTNOpts}),
      pl.handle, coords, npts, size(coords, 1), TTN_LAYOUT_AOS, dig, opts)
  end
  rc == 0 || error("ttn_digits: " * unsafe_string(ccall((:ttn_last_error, LIBTTNEVAL), Cstring, ())))
  return dig, pl.packed.site_inds
end

"""
    pin!(A::Array) / unpin!(A::Array)

Optional: page-lock a Julia array that is passed to `evaluate` many times (`ttn_host_register`, portable), so that the
library copies from / into it in place at the PCIe rate instead of staging it through its pinned ring.  Unpin before
the array is freed or resized.  Plain (pageable) arrays need neither call.
"""
function pin!(A::Array)
  rc = GC.@preserve A ccall((:ttn_host_register, LIBTTNEVAL), Cint, (Ptr{Cvoid}, UInt64), pointer(A), UInt64(sizeof(A)))
  rc == 0 || error("ttn_host_register: " * unsafe_string(ccall((:ttn_last_error, LIBTTNEVAL), Cstring, ())))
  return A
end
function unpin!(A::Array)
  rc = GC.@preserve A ccall((:ttn_host_unregister, LIBTTNEVAL), Cint, (Ptr{Cvoid},), pointer(A))
  rc == 0 || error("ttn_host_unregister: " * unsafe_string(ccall((:ttn_last_error, LIBTTNEVAL), Cstring, ())))
  return A
end
