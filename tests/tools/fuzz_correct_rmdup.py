"""TEST INFRASTRUCTURE: random-input comparison of a `siga` front-end binary with the compiled reference (oracle/_ref/siga).

    python tests/tools/fuzz_correct_rmdup.py [SEED] [CASES] [--cli PATH]

--cli defaults to build/siga_hostemu (the front-end sources on the host stand-in of the C ABI, CPU only); on a GPU box
pass siga_b200/bin/siga to run the same inputs through the CUDA library.  Mismatching cases are kept under
$TMPDIR/siga_fuzz_bad_*.  The seeded subsets in tests/test_cpu_cli.py use the same generators.
Inputs: random read sets with up to 8 % errors, FASTA or FASTQ, random -k / -x / -i / -O; compared: the file `siga correct`
writes byte for byte, the record sets of `siga rmdup`.
"""
import os, random, subprocess, sys, tempfile, shutil
REPO=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF=os.path.join(REPO,"oracle","_ref","siga"); OURS=os.path.join(REPO,"build","siga_hostemu")
if "--cli" in sys.argv:
    k=sys.argv.index("--cli"); OURS=os.path.abspath(sys.argv[k+1]); del sys.argv[k:k+2]
seed=int(sys.argv[1]) if len(sys.argv)>1 else 1
rnd=random.Random(seed)
N=int(sys.argv[2]) if len(sys.argv)>2 else 100
def rc(s): return s[::-1].translate(str.maketrans("ACGT","TGCA"))
work=tempfile.mkdtemp(); bad=0; done=0
for it in range(N):
    G=rnd.randint(60,500)
    kind=rnd.random()
    if kind<0.6: g="".join(rnd.choice("ACGT") for _ in range(G))
    elif kind<0.8:
        u="".join(rnd.choice("ACGT") for _ in range(rnd.randint(2,9))); g=(u*(G//len(u)+1))[:G]
    else: g="".join(rnd.choice("AC") if rnd.random()<0.85 else rnd.choice("GT") for _ in range(G))
    fixed=rnd.random()<0.5; L0=rnd.randint(12,min(60,G)); cov=rnd.randint(3,25)
    n=max(2,G*cov//L0)
    reads=[]
    err=rnd.choice([0.0,0.01,0.03,0.08])
    for i in range(n):
        L=L0 if fixed else rnd.randint(6,min(60,G))
        p=rnd.randrange(0,G-L+1); s=g[p:p+L]
        if rnd.random()<0.5: s=rc(s)
        s="".join(c if rnd.random()>=err else rnd.choice("ACGT") for c in s)
        reads.append(s)
    fq=rnd.random()<0.5
    d=os.path.join(work,"c%d"%it); os.makedirs(d+"/ref"); os.makedirs(d+"/ours")
    src=os.path.join(d,"in.fq" if fq else "in.fa")
    with open(src,"w") as f:
        for i,s in enumerate(reads):
            if fq:
                q="".join(rnd.choice("#+5?I") for _ in s)
                f.write("@r%d%s\n%s\n+\n%s\n"%(i,rnd.choice([""," c"]),s,q))
            else: f.write(">r%d\n%s\n"%(i,s))
    k=rnd.randint(3,min(31,L0 if fixed else 12))
    opts=["-k",str(k),"-x",str(rnd.randint(1,5)),"-i",str(rnd.randint(0,12)),"-O",str(rnd.randint(0,3))]
    ok=True
    for tool,sub in ((REF,"ref"),(OURS,"ours")):
        a=subprocess.run([tool,"index","-p","x",src],cwd=d+"/"+sub,capture_output=True,text=True)
        b=subprocess.run([tool,"correct","-p","x","-o","out.fa","-t","2"]+opts+[src],cwd=d+"/"+sub,capture_output=True,text=True)
        c=subprocess.run([tool,"rmdup","-p","x",src],cwd=d+"/"+sub,capture_output=True,text=True)
        if a.returncode or b.returncode or c.returncode:
            print(it,sub,"rc",a.returncode,b.returncode,c.returncode,(b.stderr+c.stderr)[-300:]); ok=False
    if not ok: continue
    done+=1
    A=open(d+"/ref/out.fa").read(); B=open(d+"/ours/out.fa").read()
    def recs(p):
        t=open(p).read().split(">")[1:]
        return sorted(t)
    r_ok=all(recs(d+"/ref/x.rmdup"+e)==recs(d+"/ours/x.rmdup"+e) for e in (".fa",".dups.fa"))
    if A!=B or not r_ok:
        bad+=1
        keep=os.path.join(tempfile.gettempdir(),"siga_fuzz_bad_cr_%d_%d"%(seed,it)); shutil.copytree(d,keep,dirs_exist_ok=True)
        print("MISMATCH",it,"correct" if A!=B else "", "rmdup" if not r_ok else "",opts,"fq" if fq else "fa","kept",keep)
print("compared",done,"mismatches",bad)
shutil.rmtree(work)
