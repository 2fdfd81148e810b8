"""Golden values of std::mt19937_64 + std::uniform_real_distribution<double>(-0.5, 0.5) as the
reference client draws them (client_distrib.cpp v2:231-232, seed 12345): compiled with g++ here.
    python tests/golden/make_mt_golden.py   -> tests/golden/mt19937_64.json"""
import json
import os
import subprocess
import tempfile

SRC = r'''
#include <random>
#include <cstdio>
int main(){ std::mt19937_64 gen(12345ULL); std::uniform_real_distribution<double> d(-0.5,0.5);
 for(int i=0;i<1000;i++){ double v=d(gen); if(i<8||i>=992) printf("%a\n", v);}
 std::mt19937_64 g2(42ULL); unsigned long long x=0; for(int i=0;i<700;i++) x=g2(); printf("%llu\n", x); }
'''
with tempfile.TemporaryDirectory() as d:
    open(os.path.join(d, "mt.cpp"), "w").write(SRC)
    subprocess.run(["g++", "-O1", "-o", os.path.join(d, "mt"), os.path.join(d, "mt.cpp")], check=True)
    lines = subprocess.run([os.path.join(d, "mt")], capture_output=True, text=True, check=True).stdout.split()
out = {"seed": 12345, "first8_hex": lines[:8], "last8_of_1000_hex": lines[8:16], "seed42_raw_700th": int(lines[16])}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mt19937_64.json"), "w") as f:
    json.dump(out, f, indent=1)
print(out)


# std::normal_distribution<double>(0, 1) on mt19937_64(42), as generate_random_B_block draws it
# (client_distrib.cpp v1:102-108)  -> tests/golden/normal_dist.json
SRC2 = r'''
#include <random>
#include <cstdio>
int main(){ std::mt19937_64 rng(42); std::normal_distribution<double> dist(0.0, 1.0);
 for(int i=0;i<2000;i++){ double v=dist(rng)*0.1; if(i<8||i>=1992) printf("%a\n", v);} }
'''
with tempfile.TemporaryDirectory() as d:
    open(os.path.join(d, "nd.cpp"), "w").write(SRC2)
    subprocess.run(["g++", "-O1", "-o", os.path.join(d, "nd"), os.path.join(d, "nd.cpp")], check=True)
    lines = subprocess.run([os.path.join(d, "nd")], capture_output=True, text=True, check=True).stdout.split()
out2 = {"seed": 42, "scale": 0.1, "first8_hex": lines[:8], "last8_of_2000_hex": lines[8:16]}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "normal_dist.json"), "w") as f:
    json.dump(out2, f, indent=1)
print(out2)
