# MimickB200.jl — `ccall` binding of libmimick_b200.so that keeps the two `mimick(...)` methods of
# MimickLinkedReads.jl (src/Main.jl:24 and :88) and their keyword surface verbatim.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image (SURVEY.md D8).
# It mirrors, call for call, mimick_b200/api.py, which is what the tests drive.  Output files are the
# reference's `<prefix>.R{1,2}.fq.gz`: blocked gzip (BGZF) members coded on the device and written as they
# arrive, closed by the end-of-file member (`simulate_to(...; bgzf=false)` writes plain text instead).
module MimickB200

export mimick

const LIB = get(ENV, "MIMICK_B200_LIB", joinpath(@__DIR__, "..", "mimick_b200", "libmimick_b200.so"))

# struct mmk_params (include/mimick_b200.h) — field order and types must match exactly
mutable struct Params
    struct_size::UInt32
    device::Int32
    fq_format::Int32
    bc_format::Int32
    coverage::Float64
    n_molecules::Int32
    attempts::Int32
    mol_coverage::Float64
    mol_length::Int64
    insert_length::Int32
    insert_stdev::Int32
    length_r1::Int32
    length_r2::Int32
    singletons::Float64
    circular::Int32
    mask::Int32
    seed::Int64
    strict_wrap::Int32
    force_generic::Int32
    max_chunk_bytes::UInt64
    plan_batch_groups::Int64
    err_sub_start::Float64   # optional sequencing-error model (all zero = off: the reference has none)
    err_sub_end::Float64
    err_ins::Float64
    err_del::Float64
    Params() = new()
end

struct PlanInfo
    n_groups::Int64; n_molecules::Int64; n_pairs::Int64; n_epochs::Int64; n_batches::Int64; host_ms::Float64
end
struct PlaceInfo
    first_molecule::Int64; n_molecules::Int64; first_pair::Int64; n_pairs::Int64
    r1_bytes::UInt64; r2_bytes::UInt64; gather_bytes::UInt64; host_ms::Float64
    phase_ms::NTuple{4,Float64}
end
struct Chunk
    r1_dev::Ptr{Cvoid}; r2_dev::Ptr{Cvoid}; r1_bytes::UInt64; r2_bytes::UInt64
    r1_offset::UInt64; r2_offset::UInt64; first_pair::Int64; n_pairs::Int64
    done::Int32; slot::Int32; kernel_ms::Float32
end

function check(ctx, rc)
    rc == 0 && return
    msg = ctx == C_NULL ? "" : unsafe_string(ccall((:mmk_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
    error("mimick_b200 status $rc: $msg")   # the reference raises error(...) at the same sites
end

function create(format::String; device=0, coverage=30, n_molecules=2, mol_coverage=0.2, mol_length=80000,
                insert_length=500, insert_stdev=50, read_length=[150, 150], singletons=0.70, circular=false,
                attempts=25, seed=-1, mask=false, strict_wrap=false)
    p = Params()
    ccall((:mmk_params_default, LIB), Cvoid, (Ref{Params},), p)
    bc = Ref{Int32}(0); fq = Ref{Int32}(0)
    rc = ccall((:mmk_interpret_format, LIB), Int32, (Cstring, Ref{Int32}, Ref{Int32}), format, bc, fq)
    rc == 0 || error("Barcode format $format is not recognized as one of: haplotagging, stlfr, tellseq, tenx")
    p.device = device; p.fq_format = fq[]; p.bc_format = bc[]
    p.coverage = Float64(coverage); p.n_molecules = n_molecules; p.attempts = attempts
    p.mol_coverage = mol_coverage; p.mol_length = mol_length
    p.insert_length = insert_length; p.insert_stdev = insert_stdev
    p.length_r1 = read_length[1]; p.length_r2 = length(read_length) > 1 ? read_length[2] : read_length[1]
    p.singletons = singletons; p.circular = circular; p.mask = mask; p.seed = seed; p.strict_wrap = strict_wrap
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:mmk_create, LIB), Int32, (Ref{Params}, Ref{Ptr{Cvoid}}), p, ctx)
    if rc != 0
        msg = ctx[] == C_NULL ? "" : unsafe_string(ccall((:mmk_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx[]))
        ctx[] == C_NULL || ccall((:mmk_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx[])
        error("mimick_b200 status $rc: $msg")
    end
    return ctx[]
end

# plan → place → stream every chunk to the two files.  `bgzf` (default, as the reference's BGZFCompressorStream
# writers, src/Main.jl:42-43): chunks are whole BGZF members coded on the device; the end-of-file member closes each file.
function simulate_to(ctx, r1_path::String, r2_path::String; bgzf::Bool=true)
    plan = Ref{PlanInfo}(); place = Ref{PlaceInfo}()
    check(ctx, ccall((:mmk_plan, LIB), Int32, (Ptr{Cvoid}, Ref{PlanInfo}), ctx, plan))
    check(ctx, ccall((:mmk_place, LIB), Int32, (Ptr{Cvoid}, Ref{PlaceInfo}), ctx, place))
    open(r1_path, "w") do f1
        open(r2_path, "w") do f2
            while true
                ch = Ref{Chunk}()
                if bgzf
                    check(ctx, ccall((:mmk_emit_chunk_bgzf, LIB), Int32, (Ptr{Cvoid}, Ref{Chunk}), ctx, ch))
                else
                    check(ctx, ccall((:mmk_emit_chunk, LIB), Int32, (Ptr{Cvoid}, Ref{Chunk}), ctx, ch))
                end
                c = ch[]
                if c.n_pairs > 0
                    b1 = Vector{UInt8}(undef, c.r1_bytes); b2 = Vector{UInt8}(undef, c.r2_bytes)
                    check(ctx, ccall((:mmk_copy_chunk, LIB), Int32, (Ptr{Cvoid}, Ref{Chunk}, Ptr{UInt8}, Ptr{UInt8}), ctx, ch, b1, b2))
                    write(f1, b1); write(f2, b2)
                end
                c.done != 0 && break
            end
            if bgzf
                eof = Vector{UInt8}(undef, 28)
                ccall((:mmk_bgzf_eof, LIB), Int32, (Ptr{UInt8}, UInt32), eof, 28)
                write(f1, eof); write(f2, eof)
            end
        end
    end
    return plan[], place[]
end

"""
    mimick(fasta::Vector{String}, format::String; kwargs...)  — MimickLinkedReads.jl src/Main.jl:24
"""
function mimick(fasta::Vector{String}, format::String; prefix::String="simulated/SIM", quiet::Bool=false, kw...)
    ctx = create(format; kw...)
    try
        for (i, f) in enumerate(fasta)
            check(ctx, ccall((:mmk_load_fasta, LIB), Int32, (Ptr{Cvoid}, Cstring, Int32), ctx, f, i))
        end
        mkpath(dirname(prefix))
        simulate_to(ctx, "$prefix.R1.fq.gz", "$prefix.R2.fq.gz")
    finally
        ccall((:mmk_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx)
    end
    return nothing
end

"""
    mimick(fasta::String, vcf::String, format::String; outdir, kwargs...)  — src/Main.jl:88
"""
function mimick(fasta::String, vcf::String, format::String; outdir::String="simulated", quiet::Bool=false, kw...)
    endswith(lowercase(vcf), "bcf") && error("BCF file type is not yet supported")
    v = Ref{Ptr{Cvoid}}(C_NULL)
    err = Vector{UInt8}(undef, 512)
    rc = ccall((:mmk_vcf_open, LIB), Int32, (Cstring, Ref{Ptr{Cvoid}}, Ptr{UInt8}, Csize_t), vcf, v, err, 512)
    rc == 0 || error(unsafe_string(pointer(err)))
    ctx = create(format; kw...)
    try
        check(ctx, ccall((:mmk_load_fasta, LIB), Int32, (Ptr{Cvoid}, Cstring, Int32), ctx, fasta, 1))
        mkpath(outdir)
        n = ccall((:mmk_vcf_n_samples, LIB), Int32, (Ptr{Cvoid},), v[])
        for s in 0:(n - 1)      # the reference threads over samples; one context per GPU can split this range
            name = unsafe_string(ccall((:mmk_vcf_sample_name, LIB), Cstring, (Ptr{Cvoid}, Int32), v[], s))
            check(ctx, ccall((:mmk_apply_vcf_sample, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32), ctx, v[], s))
            simulate_to(ctx, joinpath(outdir, "$name.R1.fq.gz"), joinpath(outdir, "$name.R2.fq.gz"))
        end
    finally
        ccall((:mmk_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx)
        ccall((:mmk_vcf_close, LIB), Cvoid, (Ptr{Cvoid},), v[])
    end
    return nothing
end

end # module
