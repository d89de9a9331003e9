* exercises: RANGES on E/L/G rows, RHS on the objective, rim objective, MI/FR/UP(<0)/FX/BV bounds,
* integer markers, two entries per card, OBJSENSE (ignored)
NAME tricky
OBJSENSE
    MAX
ROWS
 N  cost
 N  rimobj
 E  e1
 L  l1
 G  g1
 E  e2
 G  g2
COLUMNS
    x1  cost  1.5  e1  2.0
    x1  l1  -1.0
    x1  rimobj  9.0
    x2  cost  -2.25  e1  1.0
    x2  g1  3.0  e2  -4.5
    MARKER  'MARKER'  'INTORG'
    x3  l1  1.0  g1  1.0
    x3  g2  0.5
    MARKER  'MARKER'  'INTEND'
    x4  cost  0.125  e2  1.0
    x4  g2  -1.0  l1  2.5
    x5  g2  7.0
RHS
    rhs  cost  -3.5
    rhs  e1  4.0  l1  6.0
    rhs  g1  1.0  e2  -2.0
    rhs2  g2  100.0
RANGES
    rng  e1  2.0  l1  3.0
    rng  g1  -1.5  e2  -0.5
BOUNDS
 UP bnd  x1  -1.0
 MI bnd  x2
 FX bnd  x4  2.5
 FR bnd  x5
 UP bnd2  x5  1.0
ENDATA
