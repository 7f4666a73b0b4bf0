#!/usr/bin/env python
"""SASS instruction count per source line of one kernel (static code size): scripts/sass_lines.py <obj> <kernel substring> [top]"""
import re, subprocess, sys, tempfile, os, glob
obj, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = glob.glob(os.path.join(d, "*.cubin"))[0]
txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
cur, fn, cnt, tot = None, None, {}, 0
for line in txt.splitlines():
    if line.startswith("//---") and ".text." in line:
        fn = line
    elif "//## File" in line:
        m = re.search(r'File "([^"]+)", line (\d+)', line)
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and fn and pat in fn:
        cnt[cur] = cnt.get(cur, 0) + 1
        tot += 1
print("total SASS instructions:", tot, "=", tot * 16 // 1024, "KB")
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{v:6d}  {k[0]}:{k[1]}")
