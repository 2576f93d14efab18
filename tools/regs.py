import re,subprocess,sys
for f in sys.argv[1:]:
    print(f)
    s=open(f).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:.*\n)*?ptxas info\s+: Used (\d+) registers(.*)\n", s):
        dem=subprocess.run(['c++filt',m.group(1)],capture_output=True,text=True).stdout.strip()
        k=re.search(r'(lin_kernel<4, \d+, 0>|round_kernel<4, 0>|lin_kernel<6, 32, 1>|round_kernel<6, 1>)',dem)
        if not k: continue
        blk=s[m.start():m.end()]
        sp=re.search(r'(\d+) bytes spill stores',blk)
        print('  ',k.group(1), m.group(2), 'spill', sp.group(1) if sp else '0')
