# make_julia_goldens.jl — pins the CPU oracle to the REAL reference on a host that has Julia.
#
# UNEXECUTED HERE: neither the build image nor the GPU boxes have Julia (SURVEY.md D8), so parity of this
# repository is "unpinned" (DESIGN.md §6).  This script is the one command that pins it the day a Julia
# host with the reference's dependencies is at hand:
#
#     julia --project=<path to MimickLinkedReads.jl> baseline/make_julia_goldens.jl <reference checkout> [outdir]
#
# What it does
#   1. copies <reference>/mimick/MimickLinkedReads.jl to baseline/_ref/ (git-ignored) and applies the two
#      source patches below to the COPY (never to the checkout):
#        a. writers: BGZFWriter(open(..)) / SyncBGZFWriter(open(..)) at src/Main.jl:42-43,111-112 become plain
#           `open(.., "w")`, file names lose `.gz` — the reference cannot disable gzip (SURVEY D9), and the
#           comparison is on text anyway;
#        b. every unseeded `rand(...)` call site on the hot path gets an explicit `rng` first argument bound to
#           the module-level `const SIM_RNG = Ref{AbstractRNG}(Random.default_rng())` (call sites:
#           src/Main.jl:47,116; src/Common.jl:90; src/Breakpoints.jl:28,157,173,192,207,211,240;
#           src/ProcessFASTA.jl:32).  With SIM_RNG left alone the patched copy is the reference, seed for seed.
#   2. free-running goldens: runs the patched copy with seed = 0 on the reference's own fixtures
#      (test/fasta.fa.gz + test/test.vcf.gz, all five formats; the CI pair of haplotype FASTAs) and writes
#      the FASTQ text plus SHA-256 sums to <outdir> — true golden outputs of the reference.  The first thing to
#      read in them: D4 (is the N filter at src/Breakpoints.jl:100-104 really dead?) and D5 (does a circular run
#      produce reads longer than read_length, src/Breakpoints.jl:111-121?).
#   3. replay goldens: for every `*.variates` file in <outdir> (written by `python scripts/dump_variates.py`,
#      i.e. mmk_dump_variates of the CUDA path: one (tag, bits) record per variate in the reference's draw order,
#      SURVEY §8a-bis) runs the patched copy with SIM_RNG = ReplayRNG(file): the output must equal the CUDA
#      path's bytes.  ReplayRNG answers each primitive the samplers ask for from the next record and checks its tag:
#        randexp(rng)            <- MMK_P_NMOL / MOLSIZE / NFRAGS records (standard exponential, Float64 bits)
#        randn(rng)              <- MMK_P_INSERT records (standard normal)
#        rand(rng)               <- MMK_P_SINGLETON (and the uniform of the quantile branches)
#        rand(rng, a:b)          <- MMK_P_START / MMK_P_GAP / MMK_P_TARGET / MMK_P_MASK (index in 0:b-a)
#      Distributions.jl's truncated samplers must reach those primitives and nothing else; if a future
#      Distributions release draws differently the tag check fails loudly instead of comparing garbage.
using Random, SHA

length(ARGS) >= 1 || error("usage: make_julia_goldens.jl <reference checkout> [outdir]")
const REF = abspath(ARGS[1])
const OUT = abspath(length(ARGS) >= 2 ? ARGS[2] : joinpath(@__DIR__, "..", "tests", "golden", "julia"))
const COPY = joinpath(@__DIR__, "_ref", "MimickLinkedReads.jl")

function patched_copy()
    rm(COPY; force=true, recursive=true)
    mkpath(dirname(COPY))
    cp(joinpath(REF, "mimick", "MimickLinkedReads.jl"), COPY)
    chmod(COPY, 0o755; recursive=true)
    sub!(file, pairs...) = begin
        p = joinpath(COPY, "src", file)
        s = read(p, String)
        for (a, b) in pairs
            occursin(a, s) || error("$file: pattern not found (reference changed?): $a")
            s = replace(s, a => b)
        end
        write(p, s)
    end
    # 1a. plain writers
    sub!("Main.jl",
         r"BGZFWriter\(open\(\"\$prefix\.R1\.fq\.gz\", \"w\"\)[^\n]*\)" => "open(\"\$prefix.R1.fq\", \"w\")",
         r"BGZFWriter\(open\(\"\$prefix\.R2\.fq\.gz\", \"w\"\)[^\n]*\)" => "open(\"\$prefix.R2.fq\", \"w\")",
         "SyncBGZFWriter(open(\"\$outprefix.R1.fq.gz\", \"w\"))" => "open(\"\$outprefix.R1.fq\", \"w\")",
         "SyncBGZFWriter(open(\"\$outprefix.R2.fq.gz\", \"w\"))" => "open(\"\$outprefix.R2.fq\", \"w\")")
    # 1b. explicit rng
    sub!("MimickLinkedReads.jl", r"module MimickLinkedReads\n" => "module MimickLinkedReads\nusing Random\nconst SIM_RNG = Ref{Random.AbstractRNG}(Random.default_rng())\n")
    sub!("Main.jl", "rand(candidates, n_mol)" => "rand(SIM_RNG[], candidates, n_mol)")
    sub!("Common.jl", "round(rand(dist), digits=0)" => "round(rand(SIM_RNG[], dist), digits=0)")
    sub!("Breakpoints.jl",
         "rand(0:remaining_leeway)" => "rand(SIM_RNG[], 0:remaining_leeway)",
         "rand(params.molecule_length)" => "rand(SIM_RNG[], params.molecule_length)",
         "rand() <= params.singletons" => "rand(SIM_RNG[]) <= params.singletons",
         "rand(_exp)" => "rand(SIM_RNG[], _exp)",
         "rand(params.insert_size, n_reads)" => "rand(SIM_RNG[], params.insert_size, n_reads)",
         "rand(1:adjusted_end)" => "rand(SIM_RNG[], 1:adjusted_end)")
    sub!("ProcessFASTA.jl", "rand(ACGT)" => "rand(SIM_RNG[], ACGT)")
end

# ---- replay RNG -------------------------------------------------------------------------------------------
# record layout of mmk_variate (include/mimick_b200.h): UInt32 tag, UInt32 aux, then a Float64 (real variates) or an Int64
# (integer draws: the index in 0:n-1), both read here as their 64 bits
mutable struct ReplayRNG <: Random.AbstractRNG
    tags::Vector{UInt32}
    bits::Vector{UInt64}
    pos::Int
end
function ReplayRNG(path::String)
    raw = read(path)
    n = length(raw) ÷ 16
    tags = Vector{UInt32}(undef, n); bits = Vector{UInt64}(undef, n)
    for i in 1:n
        tags[i] = reinterpret(UInt32, raw[16i-15:16i-12])[1]
        bits[i] = reinterpret(UInt64, raw[16i-7:16i])[1]
    end
    ReplayRNG(tags, bits, 1)
end
const P_MASK, P_NMOL, P_TARGET, P_MOLSIZE, P_SINGLETON, P_NFRAGS, P_INSERT, P_START, P_GAP = UInt32.(1:9)
function next!(r::ReplayRNG, allowed)
    r.pos <= length(r.tags) || error("replay stream exhausted at record $(r.pos)")
    t = r.tags[r.pos] & 0xff
    t in allowed || error("replay record $(r.pos): tag $t, the reference asked for one of $allowed (draw order differs)")
    b = r.bits[r.pos]; r.pos += 1
    return b
end
Random.randexp(r::ReplayRNG) = reinterpret(Float64, next!(r, (P_NMOL, P_MOLSIZE, P_NFRAGS, P_INSERT)))
Random.randn(r::ReplayRNG) = reinterpret(Float64, next!(r, (P_INSERT,)))
Random.rand(r::ReplayRNG, ::Random.SamplerTrivial{Random.CloseOpen01{Float64}}) = reinterpret(Float64, next!(r, (P_SINGLETON, P_NMOL, P_MOLSIZE, P_NFRAGS, P_INSERT)))
# integer ranges, tuples and key sets: an index in 0:n-1
Random.rand(r::ReplayRNG, sp::Random.SamplerRangeNDL{U,T}) where {U,T} = (sp.a + (next!(r, (P_START, P_GAP, P_TARGET, P_MASK)) % T)) % T   # integer records carry the index itself
Random.rng_native_52(::ReplayRNG) = UInt64

function sha(path)
    open(path) do io
        bytes2hex(sha256(io))
    end
end

function main()
    patched_copy()
    push!(LOAD_PATH, dirname(COPY))
    pushfirst!(LOAD_PATH, COPY)
    M = Base.require(Main, :MimickLinkedReads)
    mkpath(OUT)
    fasta = joinpath(REF, "test", "fasta.fa.gz"); vcf = joinpath(REF, "test", "test.vcf.gz")
    sums = IOBuffer()
    for fmt in ("haplotagging", "stlfr", "tellseq", "10x", "standard")
        d = joinpath(OUT, "c1_$fmt")
        mkpath(d)
        Base.invokelatest(M.mimick, fasta, vcf, fmt; outdir=d, coverage=5, attempts=300, singletons=0.0, seed=0, quiet=true)
        for f in sort(readdir(d; join=true))
            println(sums, sha(f), "  ", relpath(f, OUT))
        end
    end
    # circular case for D5
    d = joinpath(OUT, "circular")
    mkpath(d)
    Base.invokelatest(M.mimick, [fasta], "stlfr"; prefix=joinpath(d, "SIM"), coverage=5, circular=true, attempts=300, singletons=0.0, seed=0, quiet=true)
    for f in sort(readdir(d; join=true))
        println(sums, sha(f), "  ", relpath(f, OUT))
    end
    write(joinpath(OUT, "SHA256SUMS"), take!(sums))
    # replay goldens
    for v in filter(f -> endswith(f, ".variates"), readdir(OUT; join=true))
        name = splitext(basename(v))[1]
        spec = joinpath(OUT, name * ".args.jl")   # written next to the variates: `(fastas=[...], format="...", kw=(;...))`
        isfile(spec) || (println("skip $name: no $spec"); continue)
        a = include(spec)
        M.SIM_RNG[] = ReplayRNG(v)
        pre = joinpath(OUT, name * ".replay")
        Base.invokelatest(M.mimick, a.fastas, a.format; prefix=pre, a.kw..., seed=-1, quiet=true)
        M.SIM_RNG[] = Random.default_rng()
        println(sha("$pre.R1.fq"), "  $name.replay.R1.fq")
        println(sha("$pre.R2.fq"), "  $name.replay.R2.fq")
    end
end

main()
