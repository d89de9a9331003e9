NAME
ROWS
 N  OBJ
 L  c1
 L  c2
COLUMNS
    x1        c1        1
    x1        c2        3
    x1        OBJ       -3
    x2        c1        2
    x2        c2        1
    x2        OBJ       -5
RHS
    rhs       c1        10
    rhs       c2        12
RANGES
BOUNDS
 LO bounds    x1        0
 PL bounds    x1
 LO bounds    x2        0
 PL bounds    x2
ENDATA
