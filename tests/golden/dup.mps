* duplicate (row,col) cards: values are summed; the reference then derives row pointers from the
* non-deduplicated list (src/mps_reader.cpp:1336-1355) -- mirrored bit-exactly
NAME dup
ROWS
 N  obj
 L  r1
 G  r2
 E  r3
COLUMNS
    a  obj  1.0  r1  1.0
    a  r1  0.5
    a  r2  2.0
    b  obj  2.0  r2  1.0
    b  r3  4.0
    c  r3  1.0  r1  3.0
RHS
    rhs  r1  4.0  r2  1.0
    rhs  r3  2.0
ENDATA
