"""TEST INFRASTRUCTURE: random-input comparison of a `siga` front-end binary with the compiled reference (oracle/_ref/siga).

    python tests/tools/fuzz_overlap.py [SEED] [CASES] [--cli PATH] [--multi]

--cli defaults to build/siga_hostemu (the front-end sources on the host stand-in of the C ABI, CPU only); on a GPU box
pass siga_b200/bin/siga to run the same inputs through the CUDA library.  Mismatching cases are kept under
$TMPDIR/siga_fuzz_bad_*.  The seeded subsets in tests/test_cpu_cli.py use the same generators.
Inputs: read sets from random, tandem-repeat, two-letter and repeat + palindrome genomes with duplicates and
reverse-complement duplicates, reads of 4 .. 45 bases; random -m / -x / --no-opposite-strand / -a; compared: index files
byte for byte, the ASQG as a sorted set of lines.
"""
import os, random, subprocess, sys, tempfile, shutil, gzip
REPO=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF=os.path.join(REPO,"oracle","_ref","siga"); OURS=os.path.join(REPO,"build","siga_hostemu")
if "--cli" in sys.argv:
    k=sys.argv.index("--cli"); OURS=os.path.abspath(sys.argv[k+1]); del sys.argv[k:k+2]
GPUS="--multi" in sys.argv  # --multi: a third of our overlap runs use --gpus 2 / 3 (needs that many devices, or the host stand-in)
if GPUS: sys.argv.remove("--multi")
rnd=random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 1)
N=int(sys.argv[2]) if len(sys.argv)>2 else 100
def rc(s): return s[::-1].translate(str.maketrans("ACGT","TGCA"))
def genome():
    kind=rnd.random()
    L=rnd.randint(30,400)
    if kind<0.3:
        return "".join(rnd.choice("ACGT") for _ in range(L))
    if kind<0.55:  # tandem repeats
        unit="".join(rnd.choice("ACGT") for _ in range(rnd.randint(1,7)))
        g=(unit*(L//len(unit)+1))[:L]
        g=list(g)
        for _ in range(rnd.randint(0,4)): g[rnd.randrange(L)]=rnd.choice("ACGT")
        return "".join(g)
    if kind<0.8:  # low complexity
        return "".join(rnd.choice("AC") if rnd.random()<0.9 else rnd.choice("GT") for _ in range(L))
    # repeats of a long segment + palindromes
    seg="".join(rnd.choice("ACGT") for _ in range(rnd.randint(10,40)))
    return seg+"".join(rnd.choice("ACGT") for _ in range(10))+seg+rc(seg)+seg[:len(seg)//2]
def reads_of(g):
    n=rnd.randint(2,120); fixed=rnd.random()<0.4; L0=rnd.randint(6,min(40,len(g)))
    out=[]
    for i in range(n):
        L=L0 if fixed else rnd.randint(4,min(45,len(g)))
        p=rnd.randrange(0,len(g)-L+1); s=g[p:p+L]
        if rnd.random()<0.4: s=rc(s)
        if rnd.random()<0.1 and out: s=rnd.choice(out)       # duplicate
        if rnd.random()<0.05 and out: s=rc(rnd.choice(out))  # rc duplicate
        if rnd.random()<0.1:
            s=list(s); s[rnd.randrange(len(s))]=rnd.choice("ACGT"); s="".join(s)
        out.append(s)
    return out
work=tempfile.mkdtemp(); bad=0; done=0
for it in range(N):
    rs=reads_of(genome())
    d=os.path.join(work,"c%d"%it); os.makedirs(d+"/ref"); os.makedirs(d+"/ours")
    src=os.path.join(d,"in.fa")
    names=list(range(len(rs))); 
    pre=rnd.choice(["r","x9","Z","a_"])
    open(src,"w").write("".join(">%s%d\n%s\n"%(pre,i,s) for i,s in enumerate(rs)))
    m=rnd.randint(3,16); opts=["-m",str(m)]
    if rnd.random()<0.3: opts.append("-x")
    if rnd.random()<0.25: opts.append("--no-opposite-strand")
    algo=rnd.choice(["sais","sais2"])
    gp=["--gpus",str(rnd.choice([2,3]))] if GPUS and rnd.random()<0.3 else []  # ours only: one forked worker per "GPU"
    ok=True
    for tool,sub in ((REF,"ref"),(OURS,"ours")):
        a=subprocess.run([tool,"index","-a",algo,"-p","x",src],cwd=d+"/"+sub,capture_output=True,text=True)
        b=subprocess.run([tool,"overlap","-p","x","-t","2"]+opts+(gp if tool==OURS else [])+[src],cwd=d+"/"+sub,capture_output=True,text=True)
        if a.returncode or b.returncode:
            print(it,sub,"rc",a.returncode,b.returncode,(a.stderr+b.stderr)[-300:]); ok=False
    if not ok: continue
    def lines(p): return sorted(gzip.open(p,"rt").read().splitlines())
    same=all(open(d+"/ref/x"+e,"rb").read()==open(d+"/ours/x"+e,"rb").read() for e in (".sai",".bwt",".rsai",".rbwt"))
    A=lines(d+"/ref/x.asqg.gz"); B=lines(d+"/ours/x.asqg.gz")
    done+=1
    if not same or A!=B:
        bad+=1
        keep=os.path.join(tempfile.gettempdir(),"siga_fuzz_bad_ov_%s_%d"%(sys.argv[1] if len(sys.argv)>1 else "1",it))
        shutil.copytree(d,keep,dirs_exist_ok=True)
        print("MISMATCH",it,"index same" if same else "INDEX DIFF",opts,algo,"kept",keep, "only-ref",len(set(A)-set(B)),"only-ours",len(set(B)-set(A)))
print("compared",done,"mismatches",bad)
shutil.rmtree(work)
