# ref_evaluate.jl — run the REFERENCE (ITensorNumericalAnalysis.jl, unmodified) on a network file written by
# itna_b200.save_ttn (`*.ttn.json`, layout in ../ttn_io.py) and write the values it returns.
#
#   julia --project=<ITensorNumericalAnalysis.jl checkout> ref_evaluate.jl net.ttn.json values.json [max_points]
#
# What it does: rebuilds an `ITensorNetworkFunction` from the file with the reference's own types (`IndsNetwork`,
# `RealIndexMap` / `ComplexIndexMap`, `IndsNetworkMap`, `ITensorNetwork`), then calls the reference's per-point
# `evaluate(fitn, xs, dims)` (src/itensornetworkfunction.jl:96-106: calculate_ind_values -> project ->
# scalar(alg = "bp")) on every point stored in the file, single-threaded as the reference is, and reports the
# measured points/s.  `scripts/compare_julia_reference.py` then checks those values against libttneval at 1e-12.
# This is how parity on RANDOM networks gets pinned to the real reference: Julia's RNG stream cannot be
# reproduced outside Julia, so the network travels from here to there.
#
# STATUS: written against ITensorNetworks 0.13 / ITensors 0.9 / NamedGraphs 0.6 (the reference's compat bounds)
# plus JSON.jl; NOT executed in the build environment (no Julia in the image).
using ITensorNumericalAnalysis
using ITensorNumericalAnalysis: RealIndexMap, ComplexIndexMap, IndsNetworkMap, ITensorNetworkFunction, evaluate
using ITensors: ITensors, Index, ITensor
using ITensorNetworks: ITensorNetwork, IndsNetwork
using NamedGraphs: NamedGraph
using Graphs: add_edge!
using Dictionaries: Dictionary, set!
using JSON: JSON

vertex_name(v) = length(v) > 1 ? Tuple(Int.(v)) : Int(v[1])

function load_ttn(path::AbstractString)
  doc = JSON.parsefile(path)
  @assert doc["format"] == "ttn-json-1"
  verts = [vertex_name(v) for v in doc["vertices"]]
  g = NamedGraph(verts)
  for e in doc["edges"]
    add_edge!(g, verts[Int(e[1]) + 1] => verts[Int(e[2]) + 1])
  end
  links = [Index(Int(e[3]), "Link,e$(k - 1)") for (k, e) in enumerate(doc["edges"])]
  is_cmap = doc["map"] == "complex"
  sites = Index[]
  index_digit, index_dimension, index_real = Dictionary(), Dictionary(), Dictionary()
  site_space = Dictionary(verts, [Index[] for _ in verts])
  for (k, s) in enumerate(doc["sites"])
    v = verts[Int(s["vertex"]) + 1]
    # the tags ComplexIndexMap / complex_digit_siteinds use (src/digit_inds.jl:60-70)
    tag = is_cmap ? (s["is_real"] ? "Digit,Real,S$(k - 1)" : "Digit,Imag,S$(k - 1)") : "Digit,S$(k - 1)"
    ind = Index(Int(s["dim"]), tag)
    push!(sites, ind)
    site_space[v] = vcat(site_space[v], ind)
    set!(index_digit, ind, Int(s["digit"]))
    set!(index_dimension, ind, Int(s["dimension"]))
    set!(index_real, ind, Bool(s["is_real"]))
  end
  s_net = IndsNetwork(g; site_space)
  imap = if is_cmap
    ComplexIndexMap(index_digit, index_dimension, index_real)
  else
    RealIndexMap(index_digit, index_dimension)
  end
  inm = IndsNetworkMap(s_net, imap)
  tensor_verts, ts = eltype(verts)[], ITensor[]
  for rec in doc["tensors"]
    axes = String.(rec["axes"])
    inds = [a[1] == 's' ? sites[parse(Int, a[2:end]) + 1] : links[parse(Int, a[2:end]) + 1] for a in axes]
    shape = Int.(rec["shape"])
    re = Float64.(rec["re"])
    data = haskey(rec, "im") ? complex.(re, Float64.(rec["im"])) : re
    # the file is C order (last axis fastest) over `axes`; Julia arrays are column-major
    arr = if isempty(shape)
      fill(data[1])
    else
      permutedims(reshape(data, reverse(shape)...), length(shape):-1:1)
    end
    push!(tensor_verts, verts[Int(rec["vertex"]) + 1])
    push!(ts, ITensor(arr, inds...))
  end
  tn = ITensorNetwork(tensor_verts, ts)
  fitn = ITensorNetworkFunction(tn, inm)
  dims = Int.(doc["dims"])
  points = if !haskey(doc, "points")
    nothing
  elseif is_cmap
    [[complex(Float64(z[1]), Float64(z[2])) for z in row] for row in doc["points"]]
  else
    [Float64.(row) for row in doc["points"]]
  end
  return fitn, dims, points
end

function main(args)
  path, out = args[1], args[2]
  fitn, dims, points = load_ttn(path)
  points === nothing && error("the network file carries no points")
  if length(args) >= 3
    points = points[1:min(end, parse(Int, args[3]))]
  end
  evaluate(fitn, points[1], dims)                       # compile
  vals = Vector{ComplexF64}(undef, length(points))
  t = @elapsed for (i, xs) in enumerate(points)
    vals[i] = evaluate(fitn, xs, dims)                  # the reference's own per-point path, alg = "bp"
  end
  println("reference evaluate: $(length(points)) points in $(round(t; digits=3)) s = ",
          "$(round(length(points) / t; digits=1)) points/s on $(Threads.nthreads()) Julia thread(s)")
  open(out, "w") do io
    JSON.print(io, Dict("points_per_s" => length(points) / t, "julia_threads" => Threads.nthreads(),
                        "values" => [[real(v), imag(v)] for v in vals]))
  end
end

abspath(PROGRAM_FILE) == @__FILE__() && main(ARGS)
