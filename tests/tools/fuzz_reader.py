"""TEST INFRASTRUCTURE: random-input comparison of a `siga` front-end binary with the compiled reference (oracle/_ref/siga).

    python tests/tools/fuzz_reader.py [SEED] [CASES] [--cli PATH]

--cli defaults to build/siga_hostemu (the front-end sources on the host stand-in of the C ABI, CPU only); on a GPU box
pass siga_b200/bin/siga to run the same inputs through the CUDA library.  Mismatching cases are kept under
$TMPDIR/siga_fuzz_bad_*.  The seeded subsets in tests/test_cpu_cli.py use the same generators.
Inputs: irregular FASTA / FASTQ files (bare and comment-only headers, records without lines, padded and blank lines,
CRLF, malformed FASTQ records); compared: the four index files of `siga index`.
"""
import os, random, subprocess, sys, tempfile, shutil
REPO=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF=os.path.join(REPO,"oracle","_ref","siga"); OURS=os.path.join(REPO,"build","siga_hostemu")
if "--cli" in sys.argv:
    k=sys.argv.index("--cli"); OURS=os.path.abspath(sys.argv[k+1]); del sys.argv[k:k+2]
rnd=random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 1)
def seq(n): return "".join(rnd.choice("ACGT") for _ in range(n))
def fasta():
    out=[]
    for i in range(rnd.randint(1,12)):
        k=rnd.random()
        nl=rnd.choice(["\n","\n","\r\n"])
        if k<0.55: out.append(">r%d%s%s"%(i, rnd.choice([""," c"," CR:i:3 x","\tt"]), nl))
        elif k<0.65: out.append(">"+nl)
        elif k<0.7: out.append("> "+nl)
        elif k<0.75: out.append("> lead"+nl)
        for _ in range(rnd.randint(0,3)):
            r=rnd.random()
            if r<0.7: out.append(rnd.choice([""," ","  "])+seq(rnd.randint(1,25))+rnd.choice([""," ","\t"])+nl)
            elif r<0.85: out.append(nl)
            else: out.append("  "+nl)
    s="".join(out)
    if rnd.random()<0.3: s=s.rstrip("\r\n")
    if not s.startswith(">"): s=">first\n"+s
    return s
def fastq():
    out=[]
    for i in range(rnd.randint(1,10)):
        s=seq(rnd.randint(1,25)); name="q%d%s"%(i, rnd.choice([""," c1"," BX:Z:a"]))
        nl=rnd.choice(["\n","\n","\r\n"])
        q="".join(rnd.choice("I@+>5") for _ in s)
        k=rnd.random()
        plus="+" if rnd.random()<0.6 else "+"+name
        if k<0.07: plus="-"
        if k>0.07 and k<0.12: q=q[:-1]
        if k>0.12 and k<0.16: plus="+other"
        rec=["@"+name,s,plus,q]
        if rnd.random()<0.1: rec.insert(rnd.randint(0,4),"")
        if rnd.random()<0.05: rec=rec[:rnd.randint(1,3)]
        out.append(nl.join(rec)+nl)
    return "".join(out)
bad=0
work=tempfile.mkdtemp()
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 200):
    fq=rnd.random()<0.4
    text=fastq() if fq else fasta()
    d=os.path.join(work,"c%d"%it); os.makedirs(d+"/ref"); os.makedirs(d+"/ours")
    src=os.path.join(d,"in.fq" if fq else "in.fa")
    open(src,"w",newline="").write(text)
    r1=subprocess.run([REF,"index","-p","x",src],cwd=d+"/ref",capture_output=True,text=True)
    r2=subprocess.run([OURS,"index","-p","x",src],cwd=d+"/ours",capture_output=True,text=True)
    def rd(p):
        try: return open(p,"rb").read()
        except OSError: return None
    same=all(rd(d+"/ref/x"+e)==rd(d+"/ours/x"+e) for e in (".sai",".bwt",".rsai",".rbwt"))
    refsai=rd(d+"/ref/x.sai")
    if r1.returncode!=0 or refsai is None:
        status="ref-failed(rc=%d) ours rc=%d"%(r1.returncode,r2.returncode)
        # reference crashed / failed: nothing to compare
        print(it,status); continue
    if not same or r2.returncode!=0:
        bad+=1
        print("MISMATCH",it,repr(text), "ours rc",r2.returncode, r2.stderr[-200:])
print("done, mismatches:",bad)
shutil.rmtree(work)
