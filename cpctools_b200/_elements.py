"""Chemical symbols in order of atomic number (index == Z, 'X' is the dummy element 0).

The reference takes this table from ``ase.data`` (``src/SOAPify/utils.py:5``); the product must
not depend on ase, so the public IUPAC list is carried here.
"""

chemical_symbols = (
    "X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn "
    "Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce "
    "Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn "
    "Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr Rf Db Sg Bh Hs Mt Ds Rg Cn Nh Fl "
    "Mc Lv Ts Og"
).split()
atomic_numbers = {symbol: z for z, symbol in enumerate(chemical_symbols)}
