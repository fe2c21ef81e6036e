# q' = RN(q + r*y) with y = RN(1/b), q = RN(a*y), r = a - b*q (exact, one rounding via FMA) : is it always RN(a/b)?
from fractions import Fraction as F
import random, struct, math
def rn(x: F) -> float:
    return float(x)          # Fraction -> float conversion rounds to nearest even (CPython: exact correctly rounded)
def fma(a,b,c): return rn(F(a)*F(b)+F(c))
random.seed(1)
bad=0; n=0
def rnd_mant():
    return 1.0 + random.getrandbits(52)/2**52
cases=[]
for _ in range(300000):
    a = rnd_mant()*2.0**random.randint(-20,20); b = rnd_mant()*2.0**random.randint(-20,20)
    cases.append((a,b))
# adversarial: b with mantissa all ones / near, a near b*k
for _ in range(50000):
    b = math.nextafter(2.0**random.randint(-5,5), 0.0)  # all-ones mantissa
    a = rnd_mant()*2.0**random.randint(-5,5)
    cases.append((a,b))
for _ in range(50000):
    b = rnd_mant(); k = random.randint(1,1000)
    a = rn(F(b)*k) ; a = math.nextafter(a, random.choice([0.0, 1e300]))
    cases.append((a,b))
for a,b in cases:
    y = rn(F(1)/F(b)); q = rn(F(a)*F(y)); r = fma(-b,q,a); q2 = fma(r,y,q)
    want = rn(F(a)/F(b))
    n+=1
    if q2 != want:
        bad+=1
        if bad<10: print("MISMATCH", a.hex(), b.hex(), q2.hex(), want.hex())
print(n, "cases,", bad, "mismatches")
