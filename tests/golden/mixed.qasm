OPENQASM 2.0;
qreg a[3];
qreg b[2];

h a[0];
x a[1];
y a[2];
z b[0];
cx a[0],b[1];
cz a[2],b[0];
u1(pi/8) b[1];
cu1(-pi/4) a[1],b[0];
cv a[0],a[2];
cu1(0.3) b[1],a[0];
