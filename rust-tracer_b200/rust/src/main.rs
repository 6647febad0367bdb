//! `rtrace`: the reference binary's command line over the GPU library (no clap / threadpool:
//! the image has no crates.io access).  UNBUILT here; ../host/main.cpp is the tested twin.
extern crate sphere_tracer;

use sphere_tracer::{FileOrAnyWriter, PPMStdoutRGBABufferWriter, RenderOptions, Renderer, Scene, ThreadPool};
use std::sync::Arc;
use std::{env, fs, io, path::Path, process};

fn usage_error(msg: &str) -> ! {
    eprintln!("error: {}\n\nUSAGE:\n    rtrace [OPTIONS] <output>\n\nFor more information try --help", msg);
    process::exit(1)
}

fn main() {
    // accepted for compatibility (the thread pool it sized is gone): RTRACEMAXPROCS, --num-cores
    let _nc_from_env = env::var("RTRACEMAXPROCS").ok().and_then(|v| v.parse::<usize>().ok()).unwrap_or(1);
    let (mut width, mut height, mut ssp, mut cores, mut level, mut gpus) =
        ("1024".to_string(), "1024".to_string(), "1".to_string(), "1".to_string(), "8".to_string(),
         env::var("RTRACE_GPUS").unwrap_or("1".to_string()));
    let mut output: Option<String> = None;
    let mut args = env::args().skip(1);
    while let Some(arg) = args.next() {
        let mut opt = |name: &str, slot: &mut String, args: &mut dyn Iterator<Item = String>| -> bool {
            if arg == name {
                *slot = args.next().unwrap_or_else(|| usage_error(&format!("The argument '{} <value>' requires a value", name)));
                true
            } else if arg.starts_with(&format!("{}=", name)) {
                *slot = arg[name.len() + 1..].to_string();
                true
            } else {
                false
            }
        };
        if opt("--width", &mut width, &mut args) || opt("--height", &mut height, &mut args)
            || opt("--samples-per-pixel", &mut ssp, &mut args) || opt("--num-cores", &mut cores, &mut args)
            || opt("--level", &mut level, &mut args) || opt("--gpus", &mut gpus, &mut args) {
            continue;
        }
        if arg == "--version" || arg == "-V" { println!("rtrace 0.2.0"); return; }
        if arg.starts_with("--") { usage_error(&format!("Found argument '{}' which wasn't expected", arg)); }
        if output.is_some() { usage_error(&format!("Found argument '{}' which wasn't expected", arg)); }
        output = Some(arg);
    }
    let output_file = output.unwrap_or_else(|| usage_error("The following required arguments were not provided:\n    <output>"));
    // main.rs:24-29,56-61: RTRACEMAXPROCS (default 1, unparsable -> 1); --num-cores <= 1 does not override it
    let nc_from_env: usize = std::env::var("RTRACEMAXPROCS").ok().and_then(|v| v.parse().ok()).unwrap_or(1);
    let num_cores: usize = cores.parse().unwrap();
    let pool = ThreadPool::new(if num_cores > 1 { num_cores } else { nc_from_env });

    let mut out = if output_file != "-" {
        let p = Path::new(&output_file);
        if p.extension().map(|e| e != "tga").unwrap_or(true) {
            println!("Output file '{}' must have the tga extension, e.g. {}", p.display(), p.with_extension("tga").display());
            return;
        }
        FileOrAnyWriter::FileWriter(io::BufWriter::new(fs::File::create(&p).unwrap()))
    } else {
        FileOrAnyWriter::AnyWriter(io::stdout())
    };

    let options = RenderOptions {
        width: width.parse().unwrap(),
        height: height.parse().unwrap(),
        samples_per_pixel: ssp.parse().unwrap(),
    };
    let scene = Arc::new(Scene::with_level(level.parse().unwrap(), gpus.parse().unwrap()));
    // main.rs:84-87, same call shape as the reference
    Renderer::render(&options, scene.clone(), &mut PPMStdoutRGBABufferWriter::new(true, &mut out), &pool);
    process::exit(0);
}
