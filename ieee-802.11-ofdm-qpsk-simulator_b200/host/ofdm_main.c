/* placeholder replaced below */
int main(void) { return 0; }
