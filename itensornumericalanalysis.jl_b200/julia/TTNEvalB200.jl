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

const TTN_ABI_VERSION = Int32(2)
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
end
TTNOpts(; reduce_sum=false) =
  TTNOpts(TTN_MEM_HOST, TTN_MEM_HOST, 0, reduce_sum ? 1 : 0, 0, 0.0, 0.0, 0.0f0, 0.0f0, 0, 0, C_NULL, TTN_MEM_HOST, 0, 0.0)

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
end

mutable struct TTNPlan
  handle::Ptr{Cvoid}
  packed::PackedNetwork
  function TTNPlan(packed::PackedNetwork; device::Integer=0)
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
      rc = ccall((:ttn_plan_create, LIBTTNEVAL), Cint, (Ref{TTNDesc}, Int32, Ref{Ptr{Cvoid}}),
        desc, Int32(device), h)
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
  @assert is_tree(fitn) "the batched evaluator needs a tree (cf. truncate, src/itensornetworkfunction.jl:115)"
  tn = itensornetwork(fitn)
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
  for v in vs
    i = vid[v]
    sites = collect(s[v])
    children = sort(filter(u -> parent[vid[u] + 1] == i, collect(neighbors(tn, v))); by=u -> vid[u])
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
    arr = isempty(axes) ? T[tn[v][]] : vec(array(permute(tn[v], reverse(axes)...)))
    append!(tensors, T.(arr))
    push!(tensor_ptr, length(tensors))
    for ind in sites
      pos = findfirst(==(dimension(imap, ind)), dims) - 1
      slot = cmap ? 2 * pos + (is_real(imap, ind) ? 0 : 1) : pos
      push!(site_dim, dim(ind)); push!(site_coord, slot); push!(site_digit, digit(imap, ind))
      append!(thr, [abs(index_value_to_scalar(imap, ind, k)) for k in 0:(dim(ind) - 1)])
      push!(thr_ptr, length(thr))
    end
    push!(site_ptr, length(site_dim))
  end
  return PackedNetwork(parent, link_dim, site_ptr, site_dim, site_coord, site_digit, thr_ptr, thr,
    tensor_ptr, tensors, vid[root], (cmap ? 2 : 1) * length(dims), eltype_c, cmap)
end

const _plan_cache = IdDict{Any,Any}()
function plan(fitn::ITensorNetworkFunction, dims; device=0)
  get!(() -> TTNPlan(pack(fitn, dims); device), get!(() -> Dict(), _plan_cache, fitn), (dims, device))
end

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
    evaluate(fitn, points::AbstractMatrix, dims; reduce=:none, device=0)
    evaluate(fitn, points::Vector{<:Vector}, dims; ...)

Batched `evaluate`: column `j` of `points` (or `points[j]`) holds the coordinates of point `j`
along `dims`.  Returns `Vector{Float64}` for real networks and `Vector{ComplexF64}` for complex
ones; `reduce=:sum` returns the sum over all points.  Negative or NaN coordinates raise an error
(the reference's digit loop does not terminate on them).
"""
function evaluate(fitn::ITensorNetworkFunction, points::AbstractMatrix,
  dims::Vector{<:Int}=dimensions(fitn); alg=default_contraction_alg(), reduce::Symbol=:none, device=0)
  @assert size(points, 1) == length(dims)
  pl = plan(fitn, dims; device)
  coords = coords_matrix(pl.packed, points)
  npts = size(coords, 2)
  T = pl.packed.is_complex ? ComplexF64 : Float64
  out = reduce == :sum ? T[] : Vector{T}(undef, npts)
  opts = TTNOpts(; reduce_sum=(reduce == :sum))
  GC.@preserve coords out pl begin
    rc = ccall((:ttn_evaluate, LIBTTNEVAL), Cint,
      (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Int32, Ptr{Cvoid}, Ref{TTNOpts}),
      pl.handle, coords, npts, size(coords, 1), TTN_LAYOUT_AOS,
      reduce == :sum ? C_NULL : pointer(out), opts)
  end
  rc == 0 || error("ttn_evaluate: " * unsafe_string(ccall((:ttn_last_error, LIBTTNEVAL), Cstring, ())))
  reduce == :sum && return pl.packed.is_complex ? complex(opts.sum_re, opts.sum_im) : opts.sum_re
  return out
end

function evaluate(fitn::ITensorNetworkFunction, points::Vector{<:Vector},
  dims::Vector{<:Int}=dimensions(fitn); kwargs...)
  return evaluate(fitn, reduce(hcat, points), dims; kwargs...)
end
