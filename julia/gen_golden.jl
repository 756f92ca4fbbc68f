# gen_golden.jl -- run the UNMODIFIED reference (Qaintensor.jl) on the committed seeded inputs and write
# tests/golden/reference_r02.json, the fixture that turns this repository's "oracle-relative" parity into
# reference-pinned parity (VERDICT r01, item 5).
#
#   julia --project=/path/to/Qaintensor.jl julia/gen_golden.jl [inputs.txt] [output.json]
#
# It could not be executed in the build environment (no `julia` binary, no network); tests/test_golden.py consumes
# the output whenever the file exists.  Input format: tests/golden/make_reference_inputs.py.  No JSON package is
# needed: the reader is line-oriented and the writer prints numbers and arrays only.
#
# What is recorded per item (all with the reference's own functions):
#   network, order only : perm = [t[3] for t in Qaintensor.contraction_order(net)]   (src/network2graph.jl:429-446, 476)
#   network with data   : the same perm, contract(net) before and after optimize_contraction_order!   (src/contract.jl:242-264)
#   matrix pair         : S = svd(A).S, k by the rule of src/svd.jl:29-33 at er = 1e-10, and
#                         contract_svd(T1, T2, (2, 1); er) as a flat column-major vector            (src/svd.jl:7-38)
using LinearAlgebra
using Qaintensor
using Qaintensor: contraction_order, contract

inputs = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "reference_inputs_r02.txt")
output = length(ARGS) >= 2 ? ARGS[2] : joinpath(@__DIR__, "..", "tests", "golden", "reference_r02.json")

function parse_complex(tokens::Vector{SubString{String}}, dims)
    n = prod(dims)
    @assert length(tokens) == 2n
    data = Vector{ComplexF64}(undef, n)
    for i in 1:n
        data[i] = complex(parse(Float64, tokens[2i - 1]), parse(Float64, tokens[2i]))
    end
    reshape(data, dims...)
end

jnum(x::Real) = repr(Float64(x))
jcomplex(z) = "[" * jnum(real(z)) * ", " * jnum(imag(z)) * "]"
jints(v) = "[" * join(string.(v), ", ") * "]"

networks = String[]   # JSON members
matrices = Dict{String,Matrix{ComplexF64}}()

lines = readlines(inputs)
i = 1
while i <= length(lines)
    global i
    tok = split(lines[i])
    if isempty(tok)
        i += 1
        continue
    end
    if tok[1] == "network"
        name, nt, nc, with_data = String(tok[2]), parse(Int, tok[3]), parse(Int, tok[4]), tok[5] == "1"
        tensors = Tensor[]
        for t in 1:nt
            tt = split(lines[i + t])
            @assert tt[1] == "tensor"
            r = parse(Int, tt[2])
            dims = [parse(Int, tt[2 + d]) for d in 1:r]
            if with_data
                push!(tensors, Tensor(parse_complex(tt[3 + r:end], dims)))
            else
                push!(tensors, Tensor(ones(ComplexF64, dims...)))
            end
        end
        contractions = Summation[]
        for c in 1:nc
            ct = split(lines[i + nt + c])
            @assert ct[1] == "contraction"
            t1, l1, t2, l2 = parse.(Int, ct[2:5])
            push!(contractions, Summation([t1 => l1, t2 => l2]))
        end
        i += nt + nc + 1
        net = GeneralTensorNetwork(tensors, contractions, Pair{Integer,Integer}[])
        perm = [t[3] for t in contraction_order(net)]
        member = "  \"" * name * "\": {\"perm\": " * jints(perm)
        if with_data
            before = contract(net)
            net2 = copy(net)
            optimize_contraction_order!(net2)
            after = contract(net2)
            member *= ", \"amplitude_default_order\": " * jcomplex(before[1]) * ", \"amplitude_optimized_order\": " * jcomplex(after[1])
        end
        push!(networks, member * "}")
    elseif tok[1] == "matrix"
        name, m, n = String(tok[2]), parse(Int, tok[3]), parse(Int, tok[4])
        matrices[name] = parse_complex(tok[5:end], [m, n])
        i += 1
    else
        error("unknown record: " * String(tok[1]))
    end
end

svd_members = String[]
er = 1e-10
for name in sort(collect(keys(matrices)))
    S = svd(matrices[name]).S
    # the rule of src/svd.jl:29-33, restated with the reference's own expressions
    tail = sqrt.(cumsum(reverse(S .^ 2)))
    r = findfirst(tail .> er)
    k = length(S) - r + 1
    push!(svd_members, "  \"" * name * "\": {\"S\": [" * join(jnum.(S), ", ") * "], \"er\": " * jnum(er) * ", \"k\": " * string(k) * "}")
end
if haskey(matrices, "svd_exp_T1") && haskey(matrices, "svd_exp_T2")
    T = contract_svd(Tensor(matrices["svd_exp_T1"]), Tensor(matrices["svd_exp_T2"]), (2, 1); er=er)
    push!(svd_members, "  \"contract_svd_T1_T2\": {\"er\": " * jnum(er) * ", \"data\": [" * join(jcomplex.(vec(T.data)), ", ") * "]}")
end

open(output, "w") do f
    println(f, "{")
    println(f, " \"generator\": \"julia/gen_golden.jl on Qaintensor.jl (unmodified), Julia " * string(VERSION) * "\",")
    println(f, " \"networks\": {")
    println(f, join(networks, ",\n"))
    println(f, " },")
    println(f, " \"svd\": {")
    println(f, join(svd_members, ",\n"))
    println(f, " }")
    println(f, "}")
end
println("written ", output)
