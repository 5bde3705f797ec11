"""Periodic-table lookup used by the oracle and (a separate copy) by the product.

The reference takes ``atomic_numbers`` from ``ase.data`` (``src/SOAPify/utils.py:5``);
ase is not installed here, so the table (public knowledge: IUPAC symbols in order
of atomic number, ``'X'`` = dummy element 0 like ase) is restated.
TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""

CHEMICAL_SYMBOLS = (
    "X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn "
    "Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce "
    "Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn "
    "Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr Rf Db Sg Bh Hs Mt Ds Rg Cn Nh Fl "
    "Mc Lv Ts Og"
).split()
assert len(CHEMICAL_SYMBOLS) == 119
ATOMIC_NUMBERS = {sym: z for z, sym in enumerate(CHEMICAL_SYMBOLS)}
