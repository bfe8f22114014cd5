/* placeholder translation unit: the D2Q9 oracle is added with the D2Q9 kernel */
int oracle2d_placeholder(void) { return 0; }
